"""FlatAdam / FlatEMA (lvg_dist/flat_optim.py) against torch.optim.Adam and the reference's EMA loop
(model/video_gan_lres.py:84-85,208-214): host logic on CPU tensors, the fused kernel under -m gpu."""
import copy
import math

import pytest
import torch

from lvg_dist.flat_optim import FlatAdam, FlatEMA, ema_beta_at
from lvg_dist.grad_sync import FlatGradSync


def _net(seed=0):
    torch.manual_seed(seed)
    # (batch norm first: a bias in front of it would have an exactly-zero gradient, which Adam turns into +-lr noise)
    net = torch.nn.Sequential(torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 13), torch.nn.Tanh(), torch.nn.Linear(13, 5), torch.nn.Linear(5, 3))
    return net


def _run(device, steps=5, betas=(0.0, 0.99), skip_param=False, with_sync=False, fused_sanitise=False):
    ref, ours = _net().to(device), _net().to(device)
    ref_ema, ours_ema = copy.deepcopy(ref).requires_grad_(False), copy.deepcopy(ours).requires_grad_(False)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=3e-3, betas=betas)
    sync = FlatGradSync(ours) if with_sync else None
    opt = FlatAdam(ours.parameters(), lr=3e-3, betas=betas, grad_sync=sync)
    FlatEMA(ours, ours_ema, opt)
    gen = torch.Generator().manual_seed(1)
    for step in range(steps):
        x = torch.randn(16, 7, generator=gen).to(device)
        for net in (ref, ours):
            y = net[:4](x) if (skip_param and step % 2 == 0) else net(x)          # the last layer gets no gradient on even steps
            y.square().mean().backward()
        lr = 3e-3 * min((step + 1) / 3, 1.0)
        opt_ref.param_groups[0]['lr'] = lr
        opt.param_groups[0]['lr'] = lr
        beta = ema_beta_at(step, 0.999, 2)
        gain = 0.5
        # reference tail: sync_grads arithmetic (single process), Adam, EMA over parameters and buffers
        for p in ref.parameters():
            if p.grad is not None:
                p.grad.mul_(gain)
                torch.nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)
        opt_ref.step()
        opt_ref.zero_grad(set_to_none=True)
        with torch.no_grad():
            for te, t in zip(list(ref_ema.parameters()) + list(ref_ema.buffers()), list(ref.parameters()) + list(ref.buffers())):
                if te.is_floating_point():
                    te.lerp_(t, 1.0 - beta)
                else:
                    te.copy_(t)
        if with_sync:
            sync.sync(gain, postprocess=not fused_sanitise)
            opt.step(grad_scale=sync.pending_scale if fused_sanitise else None, ema_beta=beta)
        else:
            opt.step(grad_scale=gain, ema_beta=beta)
        opt.zero_grad(set_to_none=True)
    return ref, ours, ref_ema, ours_ema, opt, opt_ref


def _close(a, b, tol=2e-6):
    return (a - b).abs().max().item() <= tol * max(b.abs().max().item(), 1e-3)


@pytest.mark.parametrize('kw', [dict(), dict(betas=(0.9, 0.999)), dict(skip_param=True, betas=(0.9, 0.99)),
                                dict(with_sync=True), dict(with_sync=True, fused_sanitise=True)])
def test_flat_adam_matches_torch_adam_cpu(kw):
    ref, ours, ref_ema, ours_ema, opt, opt_ref = _run('cpu', **kw)
    for a, b in zip(ours.parameters(), ref.parameters()):
        assert _close(a, b)
    for a, b in zip(list(ours_ema.parameters()) + list(ours_ema.buffers()), list(ref_ema.parameters()) + list(ref_ema.buffers())):
        assert _close(a.float(), b.float())
    # parameters are views into the flat buffer
    assert all(p.data_ptr() >= opt.flat_params.data_ptr() for p in ours.parameters())
    if kw.get('skip_param'):
        assert opt.steps[-1] < opt.steps[0]                                      # per-parameter step counts, as torch keeps them


def test_state_dict_roundtrip_cpu():
    _, ours, _, _, opt, _ = _run('cpu', steps=3)
    state = opt.state_dict()
    other = FlatAdam(_net().parameters(), lr=1.0, betas=(0.0, 0.99))
    other.load_state_dict(state)
    assert other.steps == opt.steps and other.param_groups[0]['lr'] == opt.param_groups[0]['lr']
    assert torch.equal(other.exp_avg_sq, opt.exp_avg_sq)


def test_ema_schedule():
    assert ema_beta_at(10 ** 9, 0.9999, 1000) == 0.9999
    assert math.isclose(ema_beta_at(0, 0.5, 0), 0.5)


@pytest.mark.gpu
@pytest.mark.parametrize('kw', [dict(), dict(betas=(0.9, 0.999)), dict(skip_param=True, betas=(0.9, 0.99)),
                                dict(with_sync=True, fused_sanitise=True)])
def test_flat_adam_matches_torch_adam_cuda(kw):
    ref, ours, ref_ema, ours_ema, opt, opt_ref = _run('cuda', **kw)
    for a, b in zip(ours.parameters(), ref.parameters()):
        assert _close(a, b, 5e-6)
    for a, b in zip(list(ours_ema.parameters()) + list(ours_ema.buffers()), list(ref_ema.parameters()) + list(ref_ema.buffers())):
        assert _close(a.float(), b.float(), 5e-6)


@pytest.mark.gpu
def test_fused_step_sanitises_like_sync_grads_cuda():
    p = torch.nn.Parameter(torch.ones(1000, device='cuda'))
    opt = FlatAdam([p], lr=0.1, betas=(0.0, 0.99))
    g = torch.randn(1000, device='cuda')
    g[3], g[4], g[5] = float('nan'), float('inf'), float('-inf')
    p.grad = g.clone()
    opt.step(grad_scale=0.5)
    ref = torch.nn.Parameter(torch.ones(1000, device='cuda'))
    ro = torch.optim.Adam([ref], lr=0.1, betas=(0.0, 0.99))
    ref.grad = torch.nan_to_num(g * 0.5, nan=0, posinf=1e5, neginf=-1e5)
    ro.step()
    assert torch.isfinite(p).all() and _close(p.detach(), ref.detach(), 5e-6)
    assert torch.equal(opt.flat_grads[:1000], ref.grad)                                 # the sanitised gradient is what stays in .grad


@pytest.mark.gpu
def test_graphed_callable_replays_a_step():
    from lvg_dist.flat_optim import GraphedCallable
    w = torch.randn(64, 64, device='cuda')
    x = torch.randn(8, 64, device='cuda')
    fn = GraphedCallable(lambda a: torch.relu(a @ w).sum(dim=1), x)
    x2 = torch.randn(8, 64, device='cuda')
    out = fn(x2)
    torch.testing.assert_close(out, torch.relu(x2 @ w).sum(dim=1))
