"""Record the operator-call trace of one LongVideoGAN network pass (authoring container only).

    python tools/trace_reference_workload.py [--reference /root/reference]

Instantiates the UNMODIFIED reference networks on CPU (random init, batch 1), wraps the
torch_utils.ops entry points and the convolution calls, runs one forward pass of each network
and writes every top-level call -- op name, tensor shapes, dtype, non-tensor arguments -- to
workloads/*.json. bench.py replays these traces through this repository's ops on synthetic
tensors (batch dimension scaled to the per-GPU batch); the reference itself does not travel
to the GPU box.

Passes recorded:
  lres_G   VideoGenerator(36x64) forward at T=160 (train_lres.py runs G at 128+32 frames)
  lres_D   VideoDiscriminator(seq_length=128, max_edge=64) forward on (1, 3, 128, 36, 64)
  sres_G   super-res VideoGenerator(144x256 <- 36x64, context 4, num_fp16_res=4) forward, T out frames
  sres_D   super-res VideoDiscriminator forward on T frames
CPU execution runs everything in fp32; the dtype each call would have on the GPU (fp16 layers)
is recorded from the module flags where the reference decides it (use_fp16).
"""
import argparse
import json
import os
import sys
import types
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    ap.add_argument('--sres-frames', type=int, default=8)
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    warnings.filterwarnings('ignore')
    sys.modules.setdefault('imageio', types.ModuleType('imageio'))

    from torch_utils.ops import bias_act, upfirdn2d, filtered_lrelu, conv2d_resample, conv2d_gradfix
    import torch.nn.functional as F

    calls = []
    depth = [0]
    fp16_hint = [False]

    def shape(t):
        return None if t is None else list(t.shape)

    def wrap(mod, name, rec):
        orig = getattr(mod, name)

        def wrapped(*a, **kw):
            if depth[0] == 0:
                entry = rec(*a, **kw)
                entry['fp16'] = bool(fp16_hint[0])
                calls.append(entry)
            depth[0] += 1
            try:
                return orig(*a, **kw)
            finally:
                depth[0] -= 1
        setattr(mod, name, wrapped)
        return orig

    def rec_bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
        return dict(op='bias_act', x=shape(x), b=b is not None, dim=dim, act=act, alpha=alpha,
                    gain=None if gain is None else float(gain), clamp=clamp)

    def rec_upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
        return dict(op='upfirdn2d', x=shape(x), f=shape(f), up=up, down=down, padding=padding,
                    flip_filter=flip_filter, gain=float(gain))

    def rec_filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=2 ** 0.5, slope=0.2, clamp=None,
                           flip_filter=False, impl='cuda'):
        return dict(op='filtered_lrelu', x=shape(x), fu=shape(fu), fd=shape(fd), b=b is not None, up=up, down=down,
                    padding=[int(p) for p in padding] if isinstance(padding, (list, tuple)) else int(padding),
                    gain=float(gain), slope=float(slope), clamp=clamp, flip_filter=flip_filter)

    def rec_conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
        return dict(op='conv2d_resample', x=shape(x), w=shape(w), f=shape(f), up=up, down=down, padding=padding,
                    groups=groups, flip_weight=flip_weight, flip_filter=flip_filter)

    def rec_conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
        return dict(op='conv2d', x=shape(input), w=shape(weight), stride=stride, padding=padding, groups=groups)

    def rec_conv3d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
        return dict(op='conv3d', x=shape(input), w=shape(weight), stride=stride, padding=padding, groups=groups)

    def rec_conv1d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
        return dict(op='conv1d', x=shape(input), w=shape(weight), stride=stride, padding=padding, groups=groups)

    wrap(bias_act, 'bias_act', rec_bias_act)
    wrap(upfirdn2d, 'upfirdn2d', rec_upfirdn2d)
    wrap(filtered_lrelu, 'filtered_lrelu', rec_filtered_lrelu)
    wrap(conv2d_resample, 'conv2d_resample', rec_conv2d_resample)
    wrap(conv2d_gradfix, 'conv2d', rec_conv2d)
    wrap(F, 'conv3d', rec_conv3d)
    wrap(F, 'conv1d', rec_conv1d)

    from model import generator_lres, discriminator_lres, generator_sres, discriminator_sres

    # the sres networks choose fp16 per layer through `use_fp16` and only apply it on CUDA; record the intent
    def hint_hook(module, _inputs):
        fp16_hint[0] = bool(getattr(module, 'use_fp16', False))
    out = {}
    torch.manual_seed(0)

    def run(name, fn):
        calls.clear()
        with torch.no_grad():
            fn()
        out[name] = list(calls)
        print(f'{name}: {len(calls)} top-level calls', {op: sum(c["op"] == op for c in calls) for op in sorted({c["op"] for c in calls})})

    G = generator_lres.VideoGenerator(out_height=36, out_width=64, num_fp16_layers=0, temporal_padding=8, temporal_emb_dim=1024)
    run('lres_G', lambda: G(1, 160))
    del G
    D = discriminator_lres.VideoDiscriminator(seq_length=128, max_edge=64, num_fp16_res=0)
    run('lres_D', lambda: D(torch.rand(1, 3, 128, 36, 64) * 2 - 1))
    del D

    T = args.sres_frames
    G = generator_sres.VideoGenerator(hr_height=144, hr_width=256, lr_height=36, lr_width=64, temporal_context=4,
                                      num_fp16_res=4, fourfeats=False)
    for m in G.modules():
        if hasattr(m, 'use_fp16'):
            m.register_forward_pre_hook(hint_hook)
    run('sres_G', lambda: G(torch.rand(1, 3, T + 8, 36, 64) * 2 - 1))
    del G
    fp16_hint[0] = False
    # the sres discriminator casts to fp16 regardless of device (discriminator_sres.py:301): trace it in fp32
    # with num_fp16_res=0 and mark the blocks the training config would run in fp16 (resolution >= 32)
    D = discriminator_sres.VideoDiscriminator(channels=3, seq_length=T, lr_height=36, lr_width=64, hr_height=144, hr_width=256,
                                              num_fp16_res=0)

    def res_hook(module, _inputs):
        fp16_hint[0] = getattr(module, 'resolution', 0) >= 32
    for m in D.modules():
        if hasattr(m, 'resolution'):
            m.register_forward_pre_hook(res_hook)
    run('sres_D', lambda: D(torch.rand(1, 3, T, 36, 64) * 2 - 1, torch.rand(1, 3, T, 144, 256) * 2 - 1))

    os.makedirs(os.path.join(ROOT, 'workloads'), exist_ok=True)
    meta = dict(source='tools/trace_reference_workload.py on the unmodified reference modules, CPU, batch 1',
                lres=dict(G='VideoGenerator(36x64) forward, T=160', D='VideoDiscriminator(seq 128, max_edge 64) forward',
                          frames_per_sample=128),
                sres=dict(G=f'VideoGenerator(144x256 <- 36x64, ctx 4) forward, T={T}', D=f'VideoDiscriminator forward, T={T}',
                          frames_per_sample=T))
    json.dump(dict(meta=meta, lres_G=out['lres_G'], lres_D=out['lres_D']), open(os.path.join(ROOT, 'workloads', 'lres_step.json'), 'w'))
    json.dump(dict(meta=meta, sres_G=out['sres_G'], sres_D=out['sres_D']), open(os.path.join(ROOT, 'workloads', 'sres_step.json'), 'w'))
    print('written workloads/lres_step.json, workloads/sres_step.json')


if __name__ == '__main__':
    main()
