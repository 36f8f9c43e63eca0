"""Filtered leaky ReLU: bias -> up-FIR -> gain*lrelu*clamp -> down-FIR (``torch_utils.ops.filtered_lrelu``).

Public entry point and gradient structure follow the reference module
(torch_utils/ops/filtered_lrelu.py:56-116 entry, :121-153 composition, :159-272
autograd function). The fused kernel keeps the up-sampled intermediate in
shared memory and stores only 2 bits per up-sampled sample (negative / clamped)
for the backward pass, which is the same operator run on dy with the roles of
the two filters swapped.
"""
import math
import warnings

import numpy as np
import torch

from .. import custom_ops
from . import bias_act
from . import upfirdn2d

_plugin = None


def _init():
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin('filtered_lrelu_plugin')
        upfirdn2d._init()
    return True


_get_filter_size = upfirdn2d._get_filter_size      # (width, height) of a [taps] / [fh, fw] filter, (1, 1) for None
_parse_padding = upfirdn2d._parse_padding          # int | [x, y] | [x0, x1, y0, y1] -> (px0, px1, py0, py1)


def _check_scalars(up, down, gain, slope, clamp):
    """Argument contract shared by both implementations (filtered_lrelu.py:121-135 of the reference)."""
    for name, v in (('up', up), ('down', down)):
        assert isinstance(v, int) and v >= 1, name
    assert gain == float(gain) and gain > 0
    assert slope == float(slope) and slope >= 0
    assert clamp is None or (clamp == float(clamp) and clamp >= 0)


def _output_extent(n_in, up, down, pad0, pad1, taps_up, taps_down):
    """Samples left after up-sampling by `up`, padding, the two FIR filters and decimation by `down`."""
    return (n_in * up + (pad0 + pad1) - (taps_up - 1) - (taps_down - 1) + (down - 1)) // down


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None,
                   flip_filter=False, impl='cuda'):
    """x: ``[N, C, H, W]`` float16/32. Per channel: add ``b[c]``; up-sample by `up` with FIR `fu`
    (zero insertion, padding w.r.t. the up-sampled image, result scaled by ``up**2``); multiply by
    `gain`, leaky ReLU with `slope`, clamp to ``[-clamp, clamp]``; down-sample by `down` with FIR `fd`.

    fu / fd: float32 ``[taps]`` (separable), ``[fh, fw]`` or None; padding: int, ``[x, y]`` or
    ``[x0, x1, y0, y1]``; flip_filter: False = convolution. Output ``[N, C, H', W']`` with
    ``W' = (W*up + px0 + px1 - (fu_w-1) - (fd_w-1) + (down-1)) // down``.
    """
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    if impl == 'cuda' and x.device.type == 'cuda' and _init():
        return _filtered_lrelu_cuda(up=up, down=down, padding=padding, gain=gain, slope=slope, clamp=clamp,
                                    flip_filter=flip_filter).apply(x, fu, fd, b, None, 0, 0)
    return _filtered_lrelu_ref(x, fu=fu, fd=fd, b=b, up=up, down=down, padding=padding, gain=gain, slope=slope,
                               clamp=clamp, flip_filter=flip_filter)


def _filtered_lrelu_ref(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None,
                        flip_filter=False):
    """Composition of bias_act and upfirdn2d (CPU tensors, ``impl='ref'``)."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    _check_scalars(up, down, gain, slope, clamp)
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.dtype == x.dtype
        assert b.ndim == 1 and b.shape[0] == x.shape[1]
    (fu_w, fu_h), (fd_w, fd_h) = _get_filter_size(fu), _get_filter_size(fd)
    px0, px1, py0, py1 = _parse_padding(padding)
    expect = [x.shape[0], x.shape[1], _output_extent(x.shape[2], up, down, py0, py1, fu_h, fd_h),
              _output_extent(x.shape[3], up, down, px0, px1, fu_w, fd_w)]
    dtype = x.dtype

    y = bias_act.bias_act(x=x, b=b)                                                            # bias
    y = upfirdn2d.upfirdn2d(x=y, f=fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    y = bias_act.bias_act(x=y, act='lrelu', alpha=slope, gain=gain, clamp=clamp)               # gain, lrelu, clamp
    y = upfirdn2d.upfirdn2d(x=y, f=fd, down=down, flip_filter=flip_filter)
    assert list(y.shape) == expect and y.dtype == dtype
    return y


class _Config:
    """One (up, down, padding, gain, slope, clamp, flip) flavour; ``.apply(x, fu, fd, b, si, sx, sy)``."""
    __slots__ = ('up', 'down', 'px0', 'px1', 'py0', 'py1', 'gain', 'slope', 'clamp', 'flip')

    def __init__(self, up, down, padding, gain, slope, clamp, flip_filter):
        _check_scalars(up, down, gain, slope, clamp)
        self.up, self.down = up, down
        self.px0, self.px1, self.py0, self.py1 = _parse_padding(padding)
        self.gain, self.slope = float(gain), float(slope)
        self.clamp = float(clamp if clamp is not None else 'inf')
        self.flip = bool(flip_filter)

    def key(self):
        return tuple(getattr(self, k) for k in self.__slots__)

    def apply(self, x, fu, fd, b, si, sx, sy):
        return _FilteredLRelu.apply(x, fu, fd, b, si, sx, sy, self)


_filtered_lrelu_cuda_cache = dict()


def _filtered_lrelu_cuda(up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None, flip_filter=False):
    cfg = _Config(up, down, padding, gain, slope, clamp, flip_filter)
    return _filtered_lrelu_cuda_cache.setdefault(cfg.key(), cfg)


def _as_kernel_filter(f, factor, device):
    """None -> exact 1x1 full filter; a separable single tap without resampling -> the full 1x1 filter f*f."""
    if f is None:
        return torch.ones([1, 1], dtype=torch.float32, device=device)
    assert 1 <= f.ndim <= 2
    if factor == 1 and f.ndim == 1 and f.shape[0] == 1:
        return f.square()[None]
    return f


def _warn_if_permuted(x):
    steps = [x.stride(d) for d in range(x.ndim) if x.size(d) > 1]
    if any(a < b for a, b in zip(steps, steps[1:])):
        warnings.warn('low-performance memory layout detected in filtered_lrelu input', RuntimeWarning)


class _FilteredLRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fu, fd, b, si, sx, sy, cfg):
        assert isinstance(x, torch.Tensor) and x.ndim == 4
        c = cfg
        fu = _as_kernel_filter(fu, c.up, x.device)
        fd = _as_kernel_filter(fd, c.down, x.device)
        si = torch.empty([0]) if si is None else si
        b = torch.zeros([x.shape[1]], dtype=x.dtype, device=x.device) if b is None else b
        # the 2-bit sign tensor is only produced when somebody will differentiate
        write_signs = (si.numel() == 0) and (x.requires_grad or b.requires_grad)
        _warn_if_permuted(x)

        y = so = None
        rc = -1
        if x.dtype in (torch.float16, torch.float32):
            # filters travel as kernel arguments (no global __constant__ state as in
            # filtered_lrelu.cu:78), so any stream may be current
            y, so, rc = _plugin.filtered_lrelu(x, fu, fd, b, si, c.up, c.down, c.px0, c.px1, c.py0, c.py1, sx, sy,
                                               c.gain, c.slope, c.clamp, c.flip, write_signs)
        if rc < 0:
            # no fused kernel for this configuration: same stages as separate native kernels,
            # still keeping only the packed signs for the backward pass
            y = x.add(b.unsqueeze(-1).unsqueeze(-1))
            y = upfirdn2d.upfirdn2d(x=y, f=fu, up=c.up, padding=[c.px0, c.px1, c.py0, c.py1], gain=c.up ** 2,
                                    flip_filter=c.flip)
            so = _plugin.filtered_lrelu_act_(y, si, sx, sy, c.gain, c.slope, c.clamp, write_signs)
            y = upfirdn2d.upfirdn2d(x=y, f=fd, down=c.down, flip_filter=c.flip)

        ctx.save_for_backward(fu, fd, (si if si.numel() else so))
        ctx.cfg = cfg
        ctx.x_shape = x.shape
        ctx.y_shape = y.shape
        ctx.s_ofs = sx, sy
        return y

    @staticmethod
    def backward(ctx, dy):
        fu, fd, si = ctx.saved_tensors
        c = ctx.cfg
        _, _, xh, xw = ctx.x_shape
        _, _, yh, yw = ctx.y_shape
        sx, sy = ctx.s_ofs
        dx = db = None
        for i in (1, 2, 4, 5, 6):
            assert not ctx.needs_input_grad[i]

        if ctx.needs_input_grad[0] or ctx.needs_input_grad[3]:
            fu_w, fu_h = fu.shape[-1], fu.shape[0]
            fd_w, fd_h = fd.shape[-1], fd.shape[0]
            pp = [(fu_w - 1) + (fd_w - 1) - c.px0,
                  xw * c.up - yw * c.down + c.px0 - (c.up - 1),
                  (fu_h - 1) + (fd_h - 1) - c.py0,
                  xh * c.up - yh * c.down + c.py0 - (c.up - 1)]
            gg = c.gain * (c.up ** 2) / (c.down ** 2)
            sx = sx - (fu_w - 1) + c.px0
            sy = sy - (fu_h - 1) + c.py0
            adj = _filtered_lrelu_cuda(up=c.down, down=c.up, padding=pp, gain=gg, slope=c.slope, clamp=None,
                                       flip_filter=(not c.flip))
            dx = adj.apply(dy, fd, fu, None, si, sx, sy)

        if ctx.needs_input_grad[3]:
            db = dx.sum([0, 2, 3])

        return dx, None, None, db, None, None, None, None
