"""Host logic of the data-parallel gradient exchange on CPU: world_size 2, gloo backend.

FlatGradSync (lvg_dist/grad_sync.py) must give what the reference's utils.sync_grads gives
(utils.py:104-125): the mean over ranks, times gain, NaN -> 0, +-inf -> +-1e5."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_sync(grads_per_rank, gain):
    """What utils.sync_grads computes, restated on plain tensors."""
    world = len(grads_per_rank)
    flat = torch.stack([torch.cat([g.flatten() for g in gs]) for gs in grads_per_rank]).sum(0)
    flat = flat / world
    flat = flat * gain
    return torch.nan_to_num(flat, nan=0, posinf=1e5, neginf=-1e5)


def _make_grads(rank):
    gen = torch.Generator().manual_seed(100 + rank)
    gs = [torch.randn(5, 3, generator=gen), torch.randn(7, generator=gen), torch.randn(2, 2, 2, generator=gen)]
    if rank == 0:
        gs[1][2] = float('nan')
        gs[0][0, 0] = float('inf')
        gs[2][1, 1, 1] = -float('inf')
    gs[0][1, 1] = 3e5 * (1 if rank == 0 else 1)      # finite but beyond the clamp after averaging
    return gs


def _worker(rank, world, port, ret):
    sys.path.insert(0, os.path.join(ROOT, 'long-video-gan_b200'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from lvg_dist.grad_sync import FlatGradSync, sync_grads
        torch.manual_seed(0)
        net = torch.nn.Module()
        net.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)),
                                         torch.nn.Parameter(torch.zeros(2, 2, 2))])
        sync = FlatGradSync(net)
        assert sync.packed().numel() == 15 + 7 + 8 and sync.flat.numel() % 64 == 0
        for p, g in zip(sync.params, _make_grads(rank)):
            p.grad.copy_(g.reshape(p.shape))                 # what backward() does: write through the view
        assert all(p.grad.data_ptr() >= sync.flat.data_ptr() for p in sync.params)
        sync.sync(gain=0.5)
        out1 = sync.packed()
        # a replaced .grad (e.g. optimizer.zero_grad(set_to_none=True) followed by backward) is folded back in
        for p, g in zip(sync.params, _make_grads(rank)):
            p.grad = g.reshape(p.shape).clone()
        sync.sync(gain=0.5)
        out2 = sync.packed()
        # functional form keeps one buffer per module
        s2 = sync_grads(net, gain=1.0)
        assert s2 is sync_grads(net, gain=1.0)
        ret[rank] = (out1, out2)
    finally:
        dist.destroy_process_group()


def test_flat_grad_sync_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    expect = _reference_sync([_make_grads(r) for r in range(world)], gain=0.5)
    for r in range(world):
        out1, out2 = ret[r]
        assert torch.equal(out1, out2)
        assert torch.allclose(out1, expect, rtol=1e-6, atol=0), (out1, expect)
    assert expect[0] == 1e5 and expect[15 + 2] == 0 and expect[-1] == -1e5 and expect[4] == 1.5e5   # finite values are not clamped


def test_single_process_is_a_plain_postprocess():
    sys.path.insert(0, os.path.join(ROOT, 'long-video-gan_b200'))
    from lvg_dist.grad_sync import FlatGradSync
    lin = torch.nn.Linear(4, 4)
    sync = FlatGradSync(lin)
    lin(torch.ones(2, 4)).sum().backward()
    assert lin.weight.grad.data_ptr() == sync.flat.data_ptr()          # backward accumulated into the flat buffer
    before = sync.flat.clone()
    sync.sync(gain=2.0)
    assert torch.allclose(sync.flat, before * 2.0)
    sync.zero_grad()
    assert float(sync.flat.abs().sum()) == 0 and lin.weight.grad.data_ptr() == sync.flat.data_ptr()


def _overlap_worker(rank, world, port, ret):
    sys.path.insert(0, os.path.join(ROOT, 'long-video-gan_b200'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from lvg_dist.grad_sync import FlatGradSync
        outs = []
        for overlap in (False, True):
            torch.manual_seed(0)
            net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                                      torch.nn.Linear(16, 3))
            extra = torch.nn.Parameter(torch.ones(5))          # never used in the loss: its bucket gets no hook call
            net.register_parameter('unused', extra)
            sync = FlatGradSync(net, overlap=overlap, buckets=3)
            gen = torch.Generator().manual_seed(10 + rank)
            for it in range(3):                                # several steps: the hook bookkeeping must re-arm
                sync.zero_grad()
                x = torch.randn(8, 6, generator=gen)
                net(x).square().mean().backward()
                if overlap and it == 1:
                    assert any(w is not None for w in sync._works)     # buckets left during the backward pass
                sync.sync(gain=0.5)
                outs.append(sync.packed())
        ret[rank] = outs
    finally:
        dist.destroy_process_group()


def test_overlapped_buckets_give_the_same_gradients_world2():
    # overlap=True starts the all-reduce of a bucket from a post-accumulate-grad hook while backward is still running;
    # the exchanged gradients must equal the single all-reduce after backward, step after step
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_overlap_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        outs = ret[r]
        plain, over = outs[:3], outs[3:]
        for a, b in zip(plain, over):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-8)
    for a, b in zip(ret[0], ret[1]):
        assert torch.equal(a, b)                                # both ranks hold the same averaged gradients


def _loop_worker(rank, world, port, ret):
    """The reference's update loop shape (video_gan_lres.py:100-176): requires_grad_(True) -> two backward passes ->
    requires_grad_(False) -> sync_grads(net, gain) -> opt.step() -> zero_grad(set_to_none=True)."""
    sys.path.insert(0, os.path.join(ROOT, 'long-video-gan_b200'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from lvg_dist import grad_sync
        res = {}
        for mode in ('reference', 'dropin', 'overlap2', 'final'):
            torch.manual_seed(0)
            net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
            net.register_parameter('unused', torch.nn.Parameter(torch.ones(5)))
            net.requires_grad_(False)                       # as between updates in the reference loop
            opt = torch.optim.Adam(net.parameters(), lr=1e-2)
            sync = None
            if mode == 'overlap2':
                sync = grad_sync.FlatGradSync(net, overlap=True, buckets=2, backwards_per_sync=2)
            elif mode == 'final':
                sync = grad_sync.FlatGradSync(net, overlap=True, buckets=2, backwards_per_sync=None)
            gen = torch.Generator().manual_seed(20 + rank)
            for it in range(3):
                net.requires_grad_(True)
                xa, xb = torch.randn(8, 6, generator=gen), torch.randn(8, 6, generator=gen)
                net(xa).square().mean().backward()
                if mode == 'final':
                    with sync.final_backward():
                        net(xb).tanh().mean().backward()
                else:
                    net(xb).tanh().mean().backward()
                net.requires_grad_(False)
                gain = None if it == 0 else 0.5
                if mode == 'reference':                     # utils.py:116-124 restated
                    ps = [p for p in net.parameters() if p.grad is not None]
                    flat = torch.cat([p.grad.flatten() for p in ps])
                    dist.all_reduce(flat)
                    flat = flat / world
                    flat = flat if gain is None else flat * gain
                    torch.nan_to_num(flat, nan=0, posinf=1e5, neginf=-1e5, out=flat)
                    for p, g in zip(ps, flat.split([p.numel() for p in ps])):
                        p.grad = g.reshape(p.shape)
                elif mode == 'dropin':
                    s = grad_sync.sync_grads(net, gain=gain)
                    assert s.packed().numel() == sum(p.numel() for p in net.parameters())     # built although requires_grad was False
                else:
                    if it == 1:
                        assert any(w is not None for w in sync._works)                 # buckets left during backward
                    sync.sync(gain=gain)
                assert net.unused.grad is None              # no gradient -> stays None (Adam skips it, as in the reference)
                opt.step()
                opt.zero_grad(set_to_none=True)
            res[mode] = torch.cat([p.detach().flatten() for p in net.parameters()])
        # a gradient arriving after its bucket left must raise, not be dropped
        net = torch.nn.Linear(4, 4)
        sync = grad_sync.FlatGradSync(net, overlap=True, buckets=1, backwards_per_sync=1)
        net(torch.ones(2, 4)).sum().backward()
        try:
            net(torch.ones(2, 4)).sum().backward()
            res['raised'] = False
        except RuntimeError as e:
            res['raised'] = 'backwards_per_sync' in str(e)
        sync.sync()
        ret[rank] = res
    finally:
        dist.destroy_process_group()


def test_sync_grads_in_the_reference_loop_shape_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_loop_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        res = ret[r]
        assert res['raised'] is True
        for mode in ('dropin', 'overlap2', 'final'):
            assert torch.allclose(res[mode], res['reference'], rtol=1e-5, atol=1e-7), mode
        assert not torch.equal(res['reference'], torch.zeros_like(res['reference']))
    for mode in ('reference', 'dropin', 'overlap2', 'final'):
        assert torch.equal(ret[0][mode], ret[1][mode]), mode         # replicas stay identical
