"""Fused multiply-add ``a * b + c`` (``torch_utils.ops.fma``, reference fma.py:15-58).

The reference has no caller for this op; it is kept for API completeness. CUDA
tensors run the broadcasting kernel in ``liblvg_ops.so``; gradients reduce the
broadcast axes back with ``_unbroadcast``.
"""
import torch

from .. import custom_ops

_plugin = None


def _init():
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin('fma_plugin')
    return True


def fma(a, b, c):  # => a * b + c
    return _FusedMultiplyAdd.apply(a, b, c)


def _mul(u, v):
    if u.device.type == 'cuda' and _init() and u.dtype in (torch.float16, torch.float32, torch.float64) \
            and max(u.ndim, v.ndim) <= 6 and v.dtype == u.dtype:
        return _plugin.fma(u, v, torch.zeros([], dtype=u.dtype, device=u.device))
    return u * v


class _FusedMultiplyAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        if a.device.type == 'cuda' and _init() and max(a.ndim, b.ndim, c.ndim) <= 6 \
                and torch.promote_types(torch.promote_types(a.dtype, b.dtype), c.dtype) in (torch.float16, torch.float32, torch.float64):
            out = _plugin.fma(a, b, c)
        else:
            out = torch.addcmul(c, a, b)
        ctx.save_for_backward(a, b)
        ctx.c_shape = c.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        a, b = ctx.saved_tensors
        da = db = dc = None
        if ctx.needs_input_grad[0]:
            da = _unbroadcast(_mul(dout, b), a.shape)
        if ctx.needs_input_grad[1]:
            db = _unbroadcast(_mul(dout, a), b.shape)
        if ctx.needs_input_grad[2]:
            dc = _unbroadcast(dout, ctx.c_shape)
        return da, db, dc


def _unbroadcast(x, shape):
    """Sum x over the axes along which `shape` was broadcast, and drop leading extra axes."""
    extra = x.ndim - len(shape)
    assert extra >= 0
    dims = [i for i in range(x.ndim) if x.shape[i] > 1 and (i < extra or shape[i - extra] == 1)]
    if dims:
        x = x.sum(dim=dims, keepdim=True)
    if extra:
        x = x.reshape(-1, *x.shape[extra + 1:])
    assert x.shape == shape
    return x
