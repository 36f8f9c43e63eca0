"""Fused bias + activation + gain + clamp (``torch_utils.ops.bias_act``).

Public surface and numerics follow the reference module
(torch_utils/ops/bias_act.py:21-31 activation table, :52-86 entry point,
:126-207 autograd functions); the arithmetic for CUDA tensors runs in
``liblvg_ops.so`` (csrc/bias_act.cu). CPU tensors and ``impl='ref'`` use the
composition of standard torch ops below, as in the reference.
"""
import math

import os

import torch

from .. import custom_ops
from ._util import AttrDict, as_dense, dense_like, is_absent

# name -> definition, default alpha/gain, kernel activation code (include/lvg_ops.h LVG_ACT_*),
# which saved tensor the gradient is expressed in, whether a second derivative exists.
activation_funcs = {
    'linear':   AttrDict(func=lambda x, **_: x,                                        def_alpha=0,   def_gain=1,            cuda_idx=1, ref='',  has_2nd_grad=False),
    'relu':     AttrDict(func=lambda x, **_: torch.nn.functional.relu(x),              def_alpha=0,   def_gain=math.sqrt(2), cuda_idx=2, ref='y', has_2nd_grad=False),
    'lrelu':    AttrDict(func=lambda x, alpha, **_: torch.nn.functional.leaky_relu(x, alpha), def_alpha=0.2, def_gain=math.sqrt(2), cuda_idx=3, ref='y', has_2nd_grad=False),
    'tanh':     AttrDict(func=lambda x, **_: torch.tanh(x),                            def_alpha=0,   def_gain=1,            cuda_idx=4, ref='y', has_2nd_grad=True),
    'sigmoid':  AttrDict(func=lambda x, **_: torch.sigmoid(x),                         def_alpha=0,   def_gain=1,            cuda_idx=5, ref='y', has_2nd_grad=True),
    'elu':      AttrDict(func=lambda x, **_: torch.nn.functional.elu(x),               def_alpha=0,   def_gain=1,            cuda_idx=6, ref='y', has_2nd_grad=True),
    'selu':     AttrDict(func=lambda x, **_: torch.nn.functional.selu(x),              def_alpha=0,   def_gain=1,            cuda_idx=7, ref='y', has_2nd_grad=True),
    'softplus': AttrDict(func=lambda x, **_: torch.nn.functional.softplus(x),          def_alpha=0,   def_gain=1,            cuda_idx=8, ref='y', has_2nd_grad=True),
    'swish':    AttrDict(func=lambda x, **_: torch.sigmoid(x) * x,                     def_alpha=0,   def_gain=math.sqrt(2), cuda_idx=9, ref='x', has_2nd_grad=True),
}

_plugin = None


def _init():
    """Bind the prebuilt plugin (called by the training scripts: train_lres.py:81)."""
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin('bias_act_plugin')
    return True


def _resolve(act, alpha, gain, clamp):
    assert clamp is None or clamp >= 0
    spec = activation_funcs[act]
    alpha = float(spec.def_alpha if alpha is None else alpha)
    gain = float(spec.def_gain if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    return spec, alpha, gain, clamp


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    """``clamp(act(x + b) * gain)`` in one pass; first and second order gradients.

    x: any shape; b: 1-D, matching ``x.shape[dim]``, or None; act: a key of
    ``activation_funcs``; alpha / gain default per activation; clamp: None or >= 0;
    impl: 'cuda' (native kernel for CUDA tensors) or 'ref'.
    """
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    if impl == 'cuda' and x.device.type == 'cuda' and _init():
        return _bias_act_cuda(dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp).apply(x, b)
    return _bias_act_ref(x=x, b=b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp)


def _bias_act_ref(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """Standard-op composition (what CPU tensors and ``impl='ref'`` get)."""
    assert isinstance(x, torch.Tensor)
    spec, alpha, gain, clamp = _resolve(act, alpha, gain, clamp)
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.ndim == 1
        assert 0 <= dim < x.ndim
        assert b.shape[0] == x.shape[dim]
        bshape = [1] * x.ndim
        bshape[dim] = -1
        x = x + b.reshape(bshape)
    x = spec.func(x, alpha=alpha)
    if gain != 1:
        x = x * gain
    if clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


class _Config:
    """Hyper-parameters of one bias_act flavour; ``.apply(x, b)`` runs it (the reference hands out
    a cached autograd.Function subclass per flavour, bias_act.py:126-139 -- same call shape)."""
    __slots__ = ('dim', 'act', 'spec', 'alpha', 'gain', 'clamp')

    def __init__(self, dim, act, alpha, gain, clamp):
        self.dim, self.act = dim, act
        self.spec, self.alpha, self.gain, self.clamp = _resolve(act, alpha, gain, clamp)

    @property
    def is_identity(self):
        return self.act == 'linear' and self.gain == 1 and self.clamp < 0

    def apply(self, x, b):
        return _BiasAct.apply(x, b, self)

    def call_plugin(self, x, b, xref, yref, dy, grad):
        return _plugin.bias_act(x, b, xref, yref, dy, grad, self.dim, self.spec.cuda_idx, self.alpha, self.gain, self.clamp)

    def reduce_bias_grad(self, dx):
        return dx.sum([i for i in range(dx.ndim) if i != self.dim])


_bias_act_cuda_cache = dict()


def _bias_act_cuda(dim=1, act='linear', alpha=None, gain=None, clamp=None):
    cfg = _Config(dim, act, alpha, gain, clamp)
    key = (dim, act, cfg.alpha, cfg.gain, cfg.clamp)
    return _bias_act_cuda_cache.setdefault(key, cfg)


def _use_codes(cfg, x):
    # relu / lrelu: the backward pass only needs "positive?" and "clamped?" per element -- 2 bits instead of re-reading y
    return (cfg.act in ('relu', 'lrelu') and x.dtype in (torch.float16, torch.float32) and x.numel() > 0
            and hasattr(_plugin, 'bias_act_fwd_codes') and os.environ.get('LVG_BIAS_ACT_CODES', '1') != '0')


def _fused_db():
    # The fused bias gradient folds per-warp partial sums with fp32 atomics: the summation ORDER varies from run to run
    # (differences of a few ulp). LVG_BIAS_ACT_FUSED_DB=0 computes db = dx.sum(...) like the reference -- bitwise reproducible.
    return os.environ.get('LVG_BIAS_ACT_FUSED_DB', '1') != '0'


def _dense_as(dy, shape, stride):
    """dy with exactly the given (dense) layout: the codes are indexed by memory offset."""
    # (a view with the right strides but a storage offset that is not 16-byte aligned -- e.g. the narrow of a dim-0 `cat`
    # backward -- is copied as well: the code-passing kernels use 128-bit accesses only)
    if tuple(dy.shape) == tuple(shape) and tuple(dy.stride()) == tuple(stride) and dy.data_ptr() % 16 == 0:
        return dy
    out = torch.empty_strided(shape, stride, dtype=dy.dtype, device=dy.device)
    out.copy_(dy)
    return out


class _BiasAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, b, cfg):
        x = as_dense(x)
        b = b.contiguous() if b is not None else None
        ctx.cfg = cfg
        ctx.has_b = b is not None
        ctx.codes_layout = None
        if (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) and _use_codes(cfg, x):
            res = _plugin.bias_act_fwd_codes(x, b, cfg.dim, cfg.spec.cuda_idx, cfg.alpha, cfg.gain, cfg.clamp)
            if res is not None:
                y, codes = res
                ctx.codes_layout = (tuple(y.shape), tuple(y.stride()))
                ctx.save_for_backward(codes)
                return y
        y = x
        if not cfg.is_identity or b is not None:
            y = cfg.call_plugin(x, b, None, None, None, 0)
        keep_x = 'x' in cfg.spec.ref or cfg.spec.has_2nd_grad
        ctx.save_for_backward(x if keep_x else None, b if keep_x else None, y if 'y' in cfg.spec.ref else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        cfg = ctx.cfg
        if ctx.codes_layout is not None:
            codes, = ctx.saved_tensors
            dy = _dense_as(dy, *ctx.codes_layout)
            need_x, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
            dx = db = None
            if need_x or need_b:
                if torch.is_grad_enabled():          # create_graph: stay differentiable in dy
                    dx = _BiasActGradCodes.apply(dy, codes, cfg)
                else:
                    dx, db = _plugin.bias_act_bwd_codes(dy, codes, cfg.dim, cfg.spec.cuda_idx, cfg.alpha, cfg.gain, cfg.clamp,
                                                        need_b and _fused_db())
            if need_b and db is None:
                db = cfg.reduce_bias_grad(dx)
            return dx, db, None
        x, b, y = ctx.saved_tensors
        like = y if y is not None else x
        if like is not None:
            dy = dense_like(dy, like)
        else:
            dy = as_dense(dy)
        need_x, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx = db = None
        if need_x or need_b:
            if cfg.is_identity:
                dx = dy
            elif (need_b and not torch.is_grad_enabled() and dy.dtype != torch.float64
                  and hasattr(_plugin, 'bias_act_grad_db') and _fused_db()):
                # plain backward (no create_graph): one kernel produces dx and the bias gradient
                fused = _plugin.bias_act_grad_db(dy, b, x, y, cfg.dim, cfg.spec.cuda_idx, cfg.alpha, cfg.gain, cfg.clamp)
                if fused is not None:
                    dx, db = fused
                else:
                    dx = _BiasActGrad.apply(dy, x, b, y, cfg)
            else:
                dx = _BiasActGrad.apply(dy, x, b, y, cfg)
        if need_b and db is None:
            db = cfg.reduce_bias_grad(dx)
        return dx, db, None


class _BiasActGradCodes(torch.autograd.Function):
    """dx = dy * gain * act'(.) from the 2-bit codes; linear in dy, so its own gradient is the same map (R1 penalty path)."""

    @staticmethod
    def forward(ctx, dy, codes, cfg):
        ctx.cfg = cfg
        ctx.layout = (tuple(dy.shape), tuple(dy.stride()))
        ctx.save_for_backward(codes)
        return _plugin.bias_act_bwd_codes(dy, codes, cfg.dim, cfg.spec.cuda_idx, cfg.alpha, cfg.gain, cfg.clamp, False)[0]

    @staticmethod
    def backward(ctx, d_dx):
        codes, = ctx.saved_tensors
        d_dy = None
        if ctx.needs_input_grad[0]:
            d_dy = _BiasActGradCodes.apply(_dense_as(d_dx, *ctx.layout), codes, ctx.cfg)
        return d_dy, None, None


class _BiasActGrad(torch.autograd.Function):
    """dx = dy * gain * act'(.), differentiable once more (R1 penalty path)."""

    @staticmethod
    def forward(ctx, dy, x, b, y, cfg):
        dx = cfg.call_plugin(dy, b, x, y, None, 1)
        ctx.cfg = cfg
        ctx.save_for_backward(dy if cfg.spec.has_2nd_grad else None, x, b, y)
        return dx

    @staticmethod
    def backward(ctx, d_dx):
        cfg = ctx.cfg
        dy, x, b, y = ctx.saved_tensors
        like = y if y is not None else x
        d_dx = dense_like(d_dx, like) if like is not None else as_dense(d_dx)
        d_dy = d_x = d_b = None
        if ctx.needs_input_grad[0]:
            d_dy = _BiasActGrad.apply(d_dx, x, b, y, cfg)
        if cfg.spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
            d_x = cfg.call_plugin(d_dx, b, x, y, dy, 2)
        if cfg.spec.has_2nd_grad and ctx.needs_input_grad[2]:
            d_b = cfg.reduce_bias_grad(d_x)
        return d_dy, d_x, d_b, None, None
