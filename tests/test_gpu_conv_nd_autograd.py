"""torch_utils/ops/conv_nd.py on the GPU: conv1d / conv2d / conv3d / conv_transpose2d through autograd (first order, and
the R1-style second order) on the tensor-core engine against torch's fp64 autograd of the same graph; the functional
proxy running conv3d / conv1d blocks shaped like the low-res networks' (generator_lres.py:83-125, discriminator_lres.py:108-213)."""
import math
import types

import pytest
import torch
import torch.nn.functional as F

from torch_utils.ops import conv_nd, conv2d_gradfix, bias_act

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rnd(shape, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(*shape, generator=g, device=DEV, dtype=torch.float64) * scale


CASES = [
    ('conv3d', (2, 24, 6, 9, 16), (40, 24, 3, 3, 3), dict(padding=(1, 1, 1))),
    ('conv3d', (1, 16, 5, 18, 32), (16, 16, 1, 3, 3), dict(padding=(0, 1, 1))),
    ('conv3d', (1, 16, 9, 8, 8), (24, 16, 5, 3, 3), dict(padding=(2, 1, 1))),
    ('conv1d', (2, 64, 16), (48, 64, 3), dict(padding=1)),
    ('conv2d', (1, 3 * 24, 20, 26), (3 * 40, 24, 3, 3), dict(padding=2, groups=3)),
    ('conv2d', (2, 32, 33, 40), (48, 32, 3, 3), dict(padding=0, stride=2)),
]


@pytest.mark.parametrize('dtype', [torch.float16, torch.float32], ids=['f16', 'f32'])
@pytest.mark.parametrize('fn,xs,ws,kw', CASES, ids=[f'{c[0]}-{i}' for i, c in enumerate(CASES)])
def test_autograd_first_and_second_order(fn, xs, ws, kw, dtype):
    x0, w0 = rnd(xs, 1).to(dtype), rnd(ws, 2, 1 / math.sqrt(math.prod(ws[1:]))).to(dtype)
    res = []
    for native in (True, False):
        dt = dtype if native else torch.float64
        x, w = x0.to(dt).requires_grad_(True), w0.to(dt).requires_grad_(True)
        y = getattr(conv_nd if native else F, fn)(x, w, None, **kw)
        v = rnd(tuple(y.shape), 3).to(dt)
        gx, gw = torch.autograd.grad((y * v).sum(), [x, w], create_graph=True)
        ggw, = torch.autograd.grad(gx.square().sum(), [w])            # R1: d |dL/dx|^2 / dw
        res.append([t.detach().double() for t in (y, gx, gw, ggw)])
    tol = 4e-3 if dtype == torch.float16 else 1e-4
    for name, a, r in zip(('y', 'dx', 'dw', 'd(|dx|^2)/dw'), *res):
        err = float((a - r).abs().max() / r.abs().max())
        assert err <= tol * (4 if name.startswith('d(') else 1), f'{name}: {err:.3e}'


@pytest.mark.parametrize('dtype', [torch.float16, torch.float32], ids=['f16', 'f32'])
def test_conv_transpose2d(dtype):
    x, w = rnd((2, 32, 9, 11), 4).to(dtype), rnd((32, 24, 3, 3), 5, 0.1).to(dtype)
    for kw in (dict(stride=2, padding=1, output_padding=1), dict(stride=1, padding=1), dict(stride=2, padding=0)):
        xg, wg = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        y = conv2d_gradfix.conv_transpose2d(xg, wg, **kw)
        r = F.conv_transpose2d(x.double().requires_grad_(True), w.double(), **kw)
        tol = 3e-3 if dtype == torch.float16 else 1e-4
        assert y.shape == r.shape and float((y.double() - r).abs().max()) <= tol * float(r.abs().max())
        gx, gw = torch.autograd.grad(y.float().square().sum(), [xg, wg])
        xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
        rx, rw = torch.autograd.grad(F.conv_transpose2d(xr, wr, **kw).square().sum(), [xr, wr])
        assert float((gx.double() - rx).abs().max()) <= 3 * tol * float(rx.abs().max())
        assert float((gw.double() - rw).abs().max()) <= 3 * tol * float(rw.abs().max())


def test_functional_proxy_runs_a_lowres_style_block():
    # a model file's view: F.conv3d -> bias_act -> F.conv3d(1x1x1), then a conv1d stack -- through the proxy
    torch.backends.cudnn.allow_tf32 = False          # the yardstick (cuDNN) in strict fp32, as train_lres.py:269-270
    torch.backends.cuda.matmul.allow_tf32 = False
    mod = types.ModuleType('fake_lres_model')
    mod.F = F
    conv_nd.install_functional(mod)
    x = rnd((2, 32, 8, 9, 16), 6).float().requires_grad_(True)
    w1, w2 = rnd((48, 32, 3, 3, 3), 7, 0.03).float().requires_grad_(True), rnd((16, 48, 1, 1, 1), 8, 0.1).float().requires_grad_(True)
    b1 = rnd((48,), 9).float().requires_grad_(True)

    def block(Fm):
        h = Fm.conv3d(x, w1, padding=(1, 1, 1))
        h = bias_act.bias_act(h, b1, act='lrelu', clamp=256)
        return Fm.conv3d(h, w2)
    y = block(mod.F)
    r = block(F)
    assert float((y - r).abs().max()) <= 5e-4 * float(r.abs().max())      # two split-precision convolutions in a row
    g = torch.autograd.grad(y.square().sum(), [x, w1, w2, b1])
    gr = torch.autograd.grad(r.square().sum(), [x, w1, w2, b1])
    for a, b in zip(g, gr):
        assert float((a - b).abs().max()) <= 1e-3 * float(b.abs().max())


def test_blurred_noise_fir_bank():
    # generator_lres.py:364-387: 128 right-aligned low-pass filters of very different lengths in a 5000-tap buffer
    import numpy as np
    import scipy.signal
    taps, widths, rows, L = 5000, 128, 6, 640
    filt = torch.zeros(widths, taps)
    for i, sr in enumerate(np.exp(np.linspace(np.log(6.0), np.log(2 * taps), widths))):
        nt = min(taps, max(3, int(np.ceil(sr / 2))))
        filt[i, -nt:] = torch.as_tensor(scipy.signal.firwin(numtaps=nt, cutoff=1.0, width=2.0, fs=float(sr)), dtype=torch.float32)
    filt = filt[:, None].to(DEV)
    noise = rnd((rows, 1, L + taps - 1), 10).float().repeat(1, widths, 1)          # einops.repeat 'n c t -> (n c) b t'
    y = conv_nd.conv1d(noise, filt, groups=widths)
    r = F.conv1d(noise.double(), filt.double(), groups=widths)
    assert y.shape == r.shape == (rows, widths, L)
    assert float((y.double() - r).abs().max()) <= 2e-5 * float(r.abs().max())
    # odd sizes, several tiles, a filter with no leading zeros
    x2, w2 = rnd((3, 5, 2500), 11).float(), rnd((5, 1, 77), 12).float()
    y2 = conv_nd.conv1d(x2, w2, groups=5)
    assert float((y2.double() - F.conv1d(x2.double(), w2.double(), groups=5)).abs().max()) <= 1e-5 * float(y2.abs().max())
    # with gradients requested the call takes the differentiable path (library), same numbers
    xg = x2.clone().requires_grad_(True)
    y3 = conv_nd.conv1d(xg, w2, groups=5)
    assert y3.requires_grad and float((y3 - y2).abs().max()) <= 1e-4 * float(y2.abs().max())


@pytest.mark.parametrize('dtype', [torch.float16, torch.float32], ids=['f16', 'f32'])
def test_fused_conv_bias_act_matches_the_two_ops(dtype):
    torch.backends.cudnn.allow_tf32 = False
    x = rnd((2, 24, 6, 9, 16), 20).to(dtype).requires_grad_(True)
    w = rnd((40, 24, 3, 3, 3), 21, 0.05).to(dtype).requires_grad_(True)
    b = rnd((40,), 22).to(dtype).requires_grad_(True)
    for act, clamp in (('lrelu', 0.875), ('linear', 256), ('lrelu', None)):          # clamps that fp16 represents exactly
        y = conv_nd.conv_bias_act(x, w, b, padding=(1, 1, 1), act=act, clamp=clamp)
        y2 = bias_act.bias_act(conv_nd.conv3d(x, w, padding=(1, 1, 1)), b, act=act, clamp=clamp)        # the two separate ops
        r = bias_act.bias_act(F.conv3d(x.double(), w.double(), padding=1), b.double(), act=act, clamp=clamp, impl='ref')
        tol = 3e-3 if dtype == torch.float16 else 1e-4
        assert float((y.double() - r).abs().max()) <= tol * float(r.abs().max())
        dy = rnd(tuple(y.shape), 23).to(dtype)
        g = torch.autograd.grad(y, [x, w, b], dy)
        g2 = torch.autograd.grad(y2, [x, w, b], dy)
        gr = torch.autograd.grad(r, [x, w, b], dy.double())
        for name, a, a2, c in zip('xwb', g, g2, gr):
            # fp16 storage decides "clamped?" / "negative?" on the rounded output, so a clamp-active fp16 case sits ~2 % (L2) from
            # the float64 graph for the separate ops as well; the fused op (one rounding less) must be at least as close
            e, e2 = float((a.double() - c.double()).norm() / c.double().norm()), float((a2.double() - c.double()).norm() / c.double().norm())
            assert e <= max(4 * tol, 1.1 * e2), f'{name}: fused {e:.3e}, separate ops {e2:.3e}'
