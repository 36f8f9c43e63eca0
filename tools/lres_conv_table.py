"""Per-signature time of every convolution of one low-res G+D training step (workloads/lres_step.json, per-GPU batch 8)
on the tensor-core engine: forward, input gradient, weight gradient, with the number of calls per step -- where the
convolution time of the headline step goes.   python tools/lres_conv_table.py > profiles/<round>_lres_conv_table.txt"""
import collections
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'long-video-gan_b200'))
from torch_utils import custom_ops  # noqa: E402

DEV = 'cuda'
_flush = None


def timeit(fn, iters=5, warmup=2):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        _flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    tr = json.load(open(os.path.join(ROOT, 'workloads', 'lres_step.json')))
    plug = custom_ops.get_plugin('convnd_plugin')
    batch = 8
    # passes per step: G forward x2 + backward x1, D forward x3 + backward x3 (update_G + update_D, video_gan_lres.py:100-176)
    mult = {'lres_G': (2, 1), 'lres_D': (3, 3)}
    tot = collections.Counter()
    print(f'# {torch.cuda.get_device_name()}  batch {batch}; ms per call, calls per step = layers x passes; TFLOP/s algorithmic')
    print(f'# {"net signature":66s} {"layers":>6s} {"fprop":>14s} {"dgrad":>14s} {"wgrad":>14s}   per-step ms (f / d / w)')
    for net in ('lres_G', 'lres_D'):
        agg = collections.OrderedDict()
        for c in tr[net]:
            if c['op'] not in ('conv3d', 'conv1d') or c['groups'] != 1:
                continue
            key = (c['op'], tuple(c['x']), tuple(c['w']), tuple(c['padding']) if isinstance(c['padding'], list) else (c['padding'],))
            agg[key] = agg.get(key, 0) + 1
        for (op, xs, ws, pad), layers in agg.items():
            xs = (batch,) + tuple(xs[1:])
            pad = tuple(pad) * (len(xs) - 2) if len(pad) == 1 else tuple(pad)
            x = torch.randn(*xs, device=DEV)
            w = torch.randn(*ws, device=DEV) / math.sqrt(math.prod(ws[1:]))
            y = plug.fprop(x, w, pad, 1)
            dy = torch.randn_like(y)
            flops = 2.0 * y.numel() * math.prod(ws[1:])
            tf = timeit(lambda: plug.fprop(x, w, pad, 1))
            td = timeit(lambda: plug.dgrad(dy, w, xs, pad, 1))
            tw = timeit(lambda: plug.wgrad(x, dy, ws, pad, 1))
            nf, nb = mult[net]
            sf, sd, sw = tf * layers * nf, td * layers * nb, tw * layers * nb
            tot['f'] += sf; tot['d'] += sd; tot['w'] += sw
            name = f'{net[5:]} {op} {"x".join(map(str, xs[1:]))} w {"x".join(map(str, ws))}'
            print(f'{name:68s} {layers:6d} {tf:7.3f} ({flops / tf / 1e9:4.0f}) {td:7.3f} ({flops / td / 1e9:4.0f}) {tw:7.3f} ({flops / tw / 1e9:4.0f})'
                  f'   {sf:7.2f} {sd:7.2f} {sw:7.2f}', flush=True)
            del x, w, y, dy
    print(f'# per step: forward {tot["f"]:.1f} ms, input gradients {tot["d"]:.1f} ms, weight gradients {tot["w"]:.1f} ms')


if __name__ == '__main__':
    main()
