"""Helper of test_chained_infer.py (subprocess on the GPU box): the UNMODIFIED reference generators (oracle/_ref/src/model,
random init) over this repository's ops -- the reference's schedule (generate.py:56-68: one segment at a time, torch.cat on
the device) against lvg_infer.chained.generate_video (segments batched, streamed to the host), eager and graph-captured."""
import sys
import time
import types
import warnings

import torch

warnings.filterwarnings('ignore')
sys.modules.setdefault('imageio', types.ModuleType('imageio'))
from model import generator_lres, generator_sres       # noqa: E402
from torch_utils.ops import conv_nd, bias_act           # noqa: E402
from lvg_infer.chained import generate_video            # noqa: E402

assert bias_act._init()
conv_nd.install_functional(generator_lres)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda')
torch.manual_seed(0)
lres_G = generator_lres.VideoGenerator(out_height=36, out_width=64, num_fp16_layers=0, temporal_padding=8, temporal_emb_dim=1024).to(dev).eval().requires_grad_(False)
sres_G = generator_sres.VideoGenerator(hr_height=144, hr_width=256, lr_height=36, lr_width=64, temporal_context=4, num_fp16_res=4,
                                       fourfeats=False).to(dev).eval().requires_grad_(False)
SEQ, SEG = 88, 16            # 6 segments (96 frames), cut to 88

with torch.no_grad():
    gen = torch.Generator('cuda').manual_seed(49)
    t0 = time.time()
    lr_len = -(-SEQ // SEG) * SEG + 2 * sres_G.temporal_context
    lr_ref = lres_G(1, lr_len, generator_emb=gen)
    ref = torch.cat(list(sres_G.sample_video_segments(lr_ref, SEG, generator_z=gen)), dim=2)[:, :, :SEQ]
    torch.cuda.synchronize()
    t_ref = time.time() - t0
    for graph in (False, True):
        gen = torch.Generator('cuda').manual_seed(49)
        t0 = time.time()
        lr, chunks = generate_video(lres_G, sres_G, SEQ, generator=gen, segment_length=SEG, segments_per_batch=4, as_uint8=False, graph=graph)
        got = torch.empty(3, SEQ, 144, 256)
        n = 0
        for first, frames in chunks:
            assert first == n
            got[:, first:first + frames.shape[1]] = frames
            n += frames.shape[1]
        t_ours = time.time() - t0
        assert n == SEQ and torch.equal(lr, lr_ref)
        a, b = got.double(), ref[0].double().cpu()
        err = float((a - b).abs().max() / b.abs().max())
        print(f'chained graph={graph}: {SEQ} frames, max rel diff vs the one-segment-at-a-time schedule {err:.2e}; '
              f'{t_ours:.2f} s (reference schedule {t_ref:.2f} s, first call includes warm-up)')
        assert torch.isfinite(a).all() and err <= 2e-3, err       # fp16 layers: batch-4 and batch-1 launches may pick different tilings
print('chained: ok')
