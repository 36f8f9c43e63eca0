"""Chained low-res -> super-res inference (lvg_infer/chained.py) against the reference's schedule
(generate.py:56-88, generator_sres.py:662-681): host logic with stand-in generators on CPU; the unmodified reference
networks on cuda under -m gpu (tests/chained_infer_run.py)."""
import os
import subprocess
import sys

import pytest
import torch

from lvg_infer.chained import generate_video, segment_windows, to_uint8

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'long-video-gan_b200')
SRC = os.path.join(ROOT, 'oracle', '_ref', 'src')


class _Lres(torch.nn.Module):
    def forward(self, batch, seq_length, generator_emb=None):
        return torch.randn(batch, 3, seq_length, 6, 8, generator=generator_emb).tanh()


class _Sres(torch.nn.Module):
    """Per-sample, per-window function with the call surface of the reference's super-res generator."""
    temporal_context = 2

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.randn(3, 3))

    def sample_latent_z(self, batch, generator=None):
        return torch.randn(batch, 5, generator=generator)

    def SG3(self, z, lr):
        c = self.temporal_context
        mid = lr[:, :, c:-c] + 0.25 * lr[:, :, :-2 * c] - 0.5 * lr[:, :, 2 * c:]            # uses the context frames
        y = torch.einsum('oc,nctHW->notHW', self.w, mid) * z[:, :1, None, None, None].tanh()
        return torch.nn.functional.interpolate(y, scale_factor=(1, 2, 2), mode='nearest')

    def sample_video_segments(self, lr_video, segment_length, generator_z=None):        # generator_sres.py:662-681
        z = self.sample_latent_z(lr_video.size(0), generator_z)
        for win in segment_windows(lr_video, segment_length, self.temporal_context):
            yield self.SG3(z, win)


@pytest.mark.parametrize('seq_length,seg,k', [(40, 8, 3), (64, 16, 8), (7, 4, 1), (48, 8, 2)])
def test_same_frames_as_the_reference_schedule(seq_length, seg, k):
    lres, sres = _Lres(), _Sres()
    # the reference's schedule (generate.py:56-68)
    gen = torch.Generator().manual_seed(5)
    lr_len = -(-seq_length // seg) * seg + 2 * sres.temporal_context
    lr_ref = lres(1, lr_len, generator_emb=gen)
    ref = torch.cat(list(sres.sample_video_segments(lr_ref, seg, generator_z=gen)), dim=2)[:, :, :seq_length]
    # ours
    gen = torch.Generator().manual_seed(5)
    lr, chunks = generate_video(lres, sres, seq_length, generator=gen, segment_length=seg, segments_per_batch=k, as_uint8=False)
    got = torch.empty(3, seq_length, *ref.shape[-2:])
    seen = 0
    for first, frames in chunks:
        assert first == seen and frames.shape[0] == 3
        got[:, first:first + frames.shape[1]] = frames                                   # consumed before the next chunk is requested
        seen += frames.shape[1]
    assert seen == seq_length and torch.equal(lr, lr_ref)
    torch.testing.assert_close(got, ref[0], rtol=1e-6, atol=1e-6)


def test_sink_and_uint8():
    lres, sres = _Lres(), _Sres()
    got = []
    lr, it = generate_video(lres, sres, 20, generator=torch.Generator().manual_seed(1), segment_length=4, segments_per_batch=2,
                            sink=lambda first, frames: got.append((first, frames.clone())))
    assert list(it) == [] and [f for f, _ in got] == [0, 8, 16] and got[-1][1].shape[1] == 4
    assert all(fr.dtype == torch.uint8 for _, fr in got)
    x = torch.tensor([-1.0, 0.0, 1.0, 3.0])
    assert to_uint8(x).tolist() == [0, 128, 255, 255]


def test_window_validation():
    with pytest.raises(ValueError):
        segment_windows(torch.zeros(1, 3, 21, 2, 2), 8, 2)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(os.path.join(SRC, 'model')), reason='oracle/_ref not staged')
def test_reference_generators_chained_on_cuda(tmp_path):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([PKG, SRC]))
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'chained_infer_run.py')], env=env, cwd=SRC,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:]
    assert 'chained: ok' in r.stdout, r.stdout[-3000:]
