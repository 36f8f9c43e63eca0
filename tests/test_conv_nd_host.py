"""Host logic of torch_utils/ops/conv_nd.py on CPU: the autograd wiring (first and second order, strides, transposed
convolution, the no_weight_gradients switch, the functional proxy) with a torch-backed stand-in for the native plugin."""
import types

import pytest
import torch
import torch.nn.functional as F

from torch_utils.ops import conv_nd, conv2d_gradfix


class TorchBackedPlugin:
    """convnd_plugin stand-in: same methods, torch arithmetic (float64 CPU). Test-only."""

    def __init__(self):
        self.calls = dict(fprop=0, dgrad=0, wgrad=0, backward=0)

    def supported(self, x, w, stride, padding, dilation, groups):
        return x.ndim in (3, 4, 5) and all(d == 1 for d in dilation)

    @staticmethod
    def _f(nd):
        return (F.conv1d, F.conv2d, F.conv3d)[nd - 1]

    @staticmethod
    def _st(nd, stride):
        return (1,) * (nd - 2) + (stride,) * min(nd, 2) if nd >= 2 else (1,)

    def fprop(self, x, w, padding, groups, stride=1, **kw):
        self.calls['fprop'] += 1
        nd = x.ndim - 2
        return self._f(nd)(x, w, None, self._st(nd, stride), padding, 1, groups)

    def dgrad(self, dy, w, x_shape, padding, groups, stride=1):
        self.calls['dgrad'] += 1
        nd = dy.ndim - 2
        x = torch.zeros(x_shape, dtype=dy.dtype, requires_grad=True)
        with torch.enable_grad():
            y = self._f(nd)(x, w, None, self._st(nd, stride), padding, 1, groups)
            return torch.autograd.grad(y, [x], dy)[0]

    def wgrad(self, x, dy, w_shape, padding, groups, stride=1):
        self.calls['wgrad'] += 1
        nd = x.ndim - 2
        w = torch.zeros(w_shape, dtype=x.dtype, requires_grad=True)
        with torch.enable_grad():
            y = self._f(nd)(x, w, None, self._st(nd, stride), padding, 1, groups)
            return torch.autograd.grad(y, [w], dy)[0]

    def backward(self, x, dy, w, padding, groups, stride=1):
        self.calls['backward'] += 1
        n = dict(self.calls)
        out = self.dgrad(dy, w, tuple(x.shape), padding, groups, stride), self.wgrad(x, dy, tuple(w.shape), padding, groups, stride)
        self.calls.update(dgrad=n['dgrad'], wgrad=n['wgrad'])
        return out


@pytest.fixture
def plug(monkeypatch):
    p = TorchBackedPlugin()
    monkeypatch.setattr(conv_nd, '_plugin', p)
    monkeypatch.setattr(conv_nd, 'enabled_for', lambda x: True)
    return p


CASES = [
    ('conv1d', (2, 6, 11), (4, 6, 3), dict(padding=1)),
    ('conv2d', (2, 6, 9, 8), (4, 3, 3, 3), dict(padding=1, groups=2)),
    ('conv2d', (2, 4, 9, 8), (5, 4, 3, 3), dict(padding=0, stride=2)),
    ('conv3d', (1, 4, 5, 6, 7), (3, 4, 3, 3, 3), dict(padding=(1, 1, 1))),
    ('conv3d', (1, 4, 5, 6, 7), (3, 4, 1, 3, 3), dict(padding=(0, 1, 1))),
]


@pytest.mark.parametrize('fn,xs,ws,kw', CASES)
def test_first_and_second_order_match_torch(plug, fn, xs, ws, kw):
    gen = torch.Generator().manual_seed(0)
    x0, w0 = torch.randn(*xs, generator=gen, dtype=torch.float64), torch.randn(*ws, generator=gen, dtype=torch.float64)
    b0 = torch.randn(ws[0], generator=gen, dtype=torch.float64)
    res = []
    for mod in (conv_nd, F):
        x, w, b = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
        y = getattr(mod, fn)(x, w, b, **kw)
        v = torch.randn(y.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
        gx, gw = torch.autograd.grad((y * v).sum(), [x, w], create_graph=True)
        pen = gx.square().sum() + gw.square().sum()                   # R1-style: gradients of gradients
        ggx, ggw, gb = torch.autograd.grad(pen + y.sum(), [x, w, b])
        res.append((y.detach(), gx.detach(), gw.detach(), ggx, ggw, gb))
    for a, r in zip(*res):
        assert torch.allclose(a, r, rtol=1e-10, atol=1e-10)
    assert plug.calls['fprop'] >= 1 and plug.calls['dgrad'] >= 1 and plug.calls['wgrad'] >= 1


def test_conv_transpose2d_is_the_input_gradient(plug):
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(2, 6, 5, 7, generator=gen, dtype=torch.float32, requires_grad=True)     # (the native path takes fp16 / fp32 only)
    w = torch.randn(6, 2, 3, 3, generator=gen, dtype=torch.float32, requires_grad=True)     # [Cin, Cout / groups, kh, kw]
    for kw in (dict(stride=2, padding=1, output_padding=1, groups=2), dict(stride=1, padding=0), dict(stride=2, padding=1)):
        y = conv_nd.conv_transpose2d(x, w, **kw)
        r = F.conv_transpose2d(x, w, **kw)
        assert torch.allclose(y, r, rtol=1e-5, atol=1e-5)
        g1 = torch.autograd.grad(y.square().sum(), [x, w])
        g2 = torch.autograd.grad(r.square().sum(), [x, w])
        for a, b in zip(g1, g2):
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-4)
    assert plug.calls['dgrad'] >= 3


def test_no_weight_gradients_switch(plug):
    x = torch.randn(1, 3, 6, 6, dtype=torch.float64, requires_grad=True)
    w = torch.randn(2, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    with conv2d_gradfix.no_weight_gradients():
        y = conv_nd.conv2d(x, w, padding=1)
        gx, gw = torch.autograd.grad(y.sum(), [x, w], allow_unused=True)
    assert gx is not None and gw is None and plug.calls['wgrad'] == 0


def test_functional_proxy_patches_only_the_convolutions(plug):
    mod = types.ModuleType('fake_model_file')
    mod.F = F
    assert conv_nd.install_functional(mod) == [mod]
    assert mod.F.pad is F.pad and mod.F.leaky_relu is F.leaky_relu
    x, w = torch.randn(1, 2, 3, 4, 5, dtype=torch.float64), torch.randn(3, 2, 1, 1, 1, dtype=torch.float64)
    n0 = plug.calls['fprop']
    assert torch.allclose(mod.F.conv3d(x, w), F.conv3d(x, w)) and plug.calls['fprop'] == n0 + 1
    other = types.ModuleType('no_F')
    assert conv_nd.install_functional(other) == []


def test_calls_outside_the_envelope_use_torch(plug):
    x, w = torch.randn(1, 2, 9, 9, dtype=torch.float64), torch.randn(2, 2, 3, 3, dtype=torch.float64)
    n0 = plug.calls['fprop']
    y = conv_nd.conv2d(x, w, padding=2, dilation=2)
    assert torch.allclose(y, F.conv2d(x, w, padding=2, dilation=2)) and plug.calls['fprop'] == n0
    assert torch.allclose(conv_nd.conv2d(x, w, padding='same'), F.conv2d(x, w, padding='same'))


@pytest.mark.parametrize('fn,xs,ws,kw', CASES)
def test_first_order_backward_is_one_fused_call(plug, fn, xs, ws, kw):
    """Both gradients in a plain backward pass = ONE plugin.backward call (dy re-tiled once); with create_graph=True the two
    gradient Functions are recorded instead (they must stay differentiable)."""
    gen = torch.Generator().manual_seed(3)
    x0, w0 = torch.randn(*xs, generator=gen, dtype=torch.float64), torch.randn(*ws, generator=gen, dtype=torch.float64)
    x, w = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    getattr(conv_nd, fn)(x, w, None, **kw).square().sum().backward()
    assert plug.calls['backward'] == 1 and plug.calls['dgrad'] == 0 and plug.calls['wgrad'] == 0
    xr, wr = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    getattr(F, fn)(xr, wr, None, **kw).square().sum().backward()
    torch.testing.assert_close(x.grad, xr.grad)
    torch.testing.assert_close(w.grad, wr.grad)
    # only one gradient wanted -> the single-gradient entry points
    x2 = x0.clone().requires_grad_(True)
    getattr(conv_nd, fn)(x2, w0, None, **kw).sum().backward()
    assert plug.calls['backward'] == 1 and plug.calls['dgrad'] == 1
    # create_graph -> separate differentiable Functions
    x3, w3 = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    gx, gw = torch.autograd.grad(getattr(conv_nd, fn)(x3, w3, None, **kw).square().sum(), [x3, w3], create_graph=True)
    assert plug.calls['backward'] == 1 and gx.requires_grad and gw.requires_grad
