"""Shared test helpers: golden fixtures, tolerances, and an oracle-backed stand-in for the
native plugins so that the host-side autograd logic can be exercised without a GPU."""
import ast
import os

import numpy as np
import torch

from oracle import oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# north_star tolerances: 1e-3 relative (fp32 activations), 1e-2 (gradients)
RTOL_ACT = 1e-3
RTOL_GRAD = 1e-2


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + '.npz'))


def cases(npz):
    return ast.literal_eval(str(npz['__cases__']))


def rel_err(got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


def assert_close(got, ref, rtol, what=''):
    if isinstance(got, torch.Tensor):
        got = got.detach().float().cpu().numpy()
    if isinstance(ref, torch.Tensor):
        ref = ref.detach().float().cpu().numpy()
    err = rel_err(got, ref)
    assert err <= rtol, f'{what}: max error relative to max |ref| = {err:.3e} > {rtol:g}'


def t(a, device='cpu', dtype=torch.float32, grad=False):
    return torch.tensor(np.asarray(a), device=device, dtype=dtype).requires_grad_(grad)


def _np(x):
    return None if x is None or x.numel() == 0 else x.detach().float().cpu().numpy()


class OracleBiasActPlugin:
    """bias_act_plugin stand-in (CPU tensors, oracle arithmetic). Test-only."""

    def bias_act(self, x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp):
        name = {v: k for k, v in orc.ACT_CODES.items()}[act]
        clamp = None if clamp < 0 else clamp
        if grad == 0:
            out = orc.bias_act(_np(x), _np(b), dim, name, alpha, gain, clamp)
        else:
            out = orc.bias_act_grad(_np(x), x=_np(xref), b=_np(b), y=_np(yref), dy=_np(dy), dim=dim, act=name,
                                    alpha=alpha, gain=gain, clamp=clamp, order=grad)
        return torch.empty_like(x).copy_(torch.from_numpy(out))   # same memory layout as x, like the real plugin


class OracleBiasActCodesPlugin(OracleBiasActPlugin):
    """Adds the relu / lrelu code-passing pair (lvg_bias_act_fwd_codes / _bwd_codes) so that the host logic that saves
    2-bit codes instead of y runs on CPU. The codes buffer is opaque to the host code; here one byte per element."""

    def __init__(self):
        self.fwd_calls = self.bwd_calls = 0

    def bias_act_fwd_codes(self, x, b, dim, act, alpha, gain, clamp):
        name = {v: k for k, v in orc.ACT_CODES.items()}[act]
        if name not in ('relu', 'lrelu'):
            return None
        self.fwd_calls += 1
        y = self.bias_act(x, b, None, None, None, 0, dim, act, alpha, gain, clamp)
        yn = y.detach().float().numpy()
        inv_gain = 1.0 / gain if gain != 0 else 0.0
        codes = (~(yn * np.float32(inv_gain) > 0)).astype(np.uint8) | ((clamp >= 0) & ~(np.abs(yn) < clamp)).astype(np.uint8) * 2
        return y, torch.from_numpy(np.ascontiguousarray(codes.reshape(-1)))

    def bias_act_bwd_codes(self, dy, codes, dim, act, alpha, gain, clamp, want_db):
        name = {v: k for k, v in orc.ACT_CODES.items()}[act]
        self.bwd_calls += 1
        c = codes.numpy().reshape(dy.shape)          # test tensors are contiguous: memory order == index order
        d = dy.detach().float().numpy()
        slope = alpha if name == 'lrelu' else 0.0
        dx = np.where(c & 1, d * np.float32(slope), d) * np.float32(gain)
        dx = np.where(c & 2, 0, dx).astype(np.float32)
        out = torch.empty_like(dy).copy_(torch.from_numpy(dx))
        db = None
        if want_db:
            db = out.float().sum([i for i in range(out.ndim) if i != dim]).to(dy.dtype)
        return out, db


class OracleUpfirdn2dPlugin:
    def upfirdn2d(self, x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
        out = orc.upfirdn2d(_np(x), f.numpy(), [upx, upy], [downx, downy], [padx0, padx1, pady0, pady1], flip, gain)
        return torch.from_numpy(out).to(x.dtype)


class OracleFilteredLReluPlugin:
    def __init__(self, fused=True):
        self.fused = fused

    def filtered_lrelu(self, x, fu, fd, b, si, up, down, px0, px1, py0, py1, sx, sy, gain, slope, clamp, flip, write_signs):
        if not self.fused:
            return None, None, -1
        s_in = None if si.numel() == 0 else si.numpy()
        res = orc.filtered_lrelu(_np(x), fu.numpy(), fd.numpy(), _np(b), up, down, [px0, px1, py0, py1], gain, slope,
                                 None if clamp == float('inf') else clamp, flip, signs_in=s_in, sx=sx, sy=sy,
                                 return_signs=bool(write_signs))
        y, so = res if write_signs else (res, None)
        return torch.from_numpy(y).to(x.dtype), (torch.from_numpy(so) if so is not None else None), 0

    def filtered_lrelu_act_(self, x, si, sx, sy, gain, slope, clamp, write_signs):
        # in-place activation on the up-sampled tensor == filtered_lrelu with 1x1 filters and no resampling
        s_in = None if si.numel() == 0 else si.numpy()
        one = np.ones([1, 1], np.float32)
        res = orc.filtered_lrelu(_np(x), one, one, None, 1, 1, 0, gain, slope, None if clamp == float('inf') else clamp,
                                 False, signs_in=s_in, sx=sx, sy=sy, return_signs=bool(write_signs))
        y, so = res if write_signs else (res, None)
        x.copy_(torch.from_numpy(y).to(x.dtype))
        return torch.from_numpy(so) if so is not None else None
