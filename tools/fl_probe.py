"""Runs the three fused filtered_lrelu configurations of the super-res generator once each (for ncu captures / timing):
python tools/fl_probe.py [iters]"""
import math
import os
import sys

import scipy.signal
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'long-video-gan_b200'))
from torch_utils.ops import filtered_lrelu  # noqa: E402

DEV = 'cuda'


def kaiser(taps, scale):
    return torch.tensor(scipy.signal.firwin(numtaps=taps, cutoff=0.5, width=0.6, fs=2.0 * scale), dtype=torch.float32, device=DEV)


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    cases = [('L12 up2/down2', (64, 128, 166, 278), 2, 2, [9, 8, 9, 8]), ('L10 up4/down2', (64, 256, 94, 150), 4, 2, [-6, -9, -6, -9])]
    for name, shape, up, down, pad in cases:
        fu, fd = kaiser(6 * up, up), kaiser(6 * down, down)
        x = torch.randn(*shape, device=DEV, dtype=torch.float16).requires_grad_(True)
        b = torch.randn(shape[1], device=DEV, dtype=torch.float16).requires_grad_(True)
        for _ in range(iters):
            y = filtered_lrelu.filtered_lrelu(x, fu, fd, b, up=up, down=down, padding=pad, gain=math.sqrt(2), slope=0.2, clamp=256)
            dy = torch.randn_like(y)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            torch.autograd.grad(y, [x, b], dy)
            ev[1].record()
            torch.cuda.synchronize()
            print(name, 'backward ms', ev[0].elapsed_time(ev[1]))


if __name__ == '__main__':
    main()
