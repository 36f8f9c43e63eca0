"""Times the fused update tail (lvg_adam_step: sanitise + Adam + EMA over flat buffers) at the low-res networks' parameter
counts against torch.optim.Adam (foreach) + the reference's EMA loop on the same number of elements spread over ~300 tensors.
    python tools/bench_optim.py > profiles/<round>_optim.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'long-video-gan_b200'))
from lvg_dist.flat_optim import FlatAdam  # noqa: E402


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    dev = 'cuda'
    print(f'# {torch.cuda.get_device_name()}; ms per update tail; GB/s = bytes moved by the fused kernel / time')
    for name, n, ema in (('G 83.2 M parameters + EMA', 83_200_000, True), ('D 46.4 M parameters', 46_400_000, False)):
        nt = 300
        sizes = [n // nt] * nt
        sizes[-1] += n - sum(sizes)
        ps = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.02) for s in sizes]
        for p in ps:
            p.grad = torch.randn_like(p) * 1e-3
        es = [p.detach().clone() for p in ps]
        ref = torch.optim.Adam(ps, lr=3e-3, betas=(0.0, 0.99))

        def ref_tail():
            grads = [p.grad for p in ps]
            torch._foreach_mul_(grads, 0.5)                       # utils.sync_grads: scale + nan_to_num on the flat copy (two passes)
            for g in grads[:1]:
                torch.nan_to_num(g, nan=0, posinf=1e5, neginf=-1e5, out=g)
            ref.step()
            if ema:
                with torch.no_grad():
                    for e, p in zip(es, ps):                     # video_gan_lres.py:213-214
                        e.lerp_(p, 0.001)
        t_ref = timeit(ref_tail)
        qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
        opt = FlatAdam(qs, lr=3e-3, betas=(0.0, 0.99))
        opt.flat_grads.normal_(0, 1e-3)
        for p, v in zip(opt.params, opt._grad_views):
            p.grad = v
        if ema:
            class _E:
                flat_params = torch.zeros_like(opt.flat_params)
                lerp_range = staticmethod(lambda a, b, beta: None)
                update_buffers = staticmethod(lambda beta: None)
            opt._ema = _E
        t = timeit(lambda: opt.step(grad_scale=0.5, ema_beta=0.999 if ema else None))
        gb = n * (28 + (8 if ema else 0)) / 1e9
        print(f'{name:28s} fused {t:7.3f} ms ({gb / t * 1e3:6.0f} GB/s)   torch.optim.Adam(foreach) + per-tensor EMA {t_ref:7.3f} ms   x{t_ref / t:.1f}')


if __name__ == '__main__':
    main()
