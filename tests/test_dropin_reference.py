"""Drop-in rule: with `long-video-gan_b200/` ahead of a LongVideoGAN checkout on PYTHONPATH the
reference's own `model/*` code imports THIS repository's `torch_utils.ops` while `torch_utils.misc`, `dnnlib`, ... still
come from the checkout -- and the networks compute the same thing. Runs the unmodified reference generators and
discriminator on CPU (where both resolve to compositions of standard torch ops) twice, once per package resolution, in
subprocesses, and compares the outputs. Only where a checkout is present (the authoring container: /root/reference)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get('LVG_REFERENCE_CHECKOUT', '/root/reference')

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, 'model')), reason='no LongVideoGAN checkout available')


def _run(pythonpath, out_file):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(pythonpath), CUDA_VISIBLE_DEVICES='')
    subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'dropin_reference_run.py'), out_file], check=True, env=env,
                   cwd=REFERENCE, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=1200)
    return torch.load(out_file)


def test_reference_networks_run_unchanged_on_the_dropped_in_ops(tmp_path):
    pkg = os.path.join(ROOT, 'long-video-gan_b200')
    theirs = _run([REFERENCE], str(tmp_path / 'ref.pt'))
    ours = _run([pkg, REFERENCE], str(tmp_path / 'ours.pt'))
    # module resolution: ops from here, the rest of torch_utils and dnnlib from the checkout
    assert ours['where']['bias_act'].startswith(pkg) and ours['where']['upfirdn2d'].startswith(pkg)
    assert ours['where']['misc'].startswith(REFERENCE) and ours['where']['dnnlib'].startswith(REFERENCE)
    assert theirs['where']['bias_act'].startswith(REFERENCE)
    # same networks, same seeds: low-res generator (32 frames 36x64), low-res discriminator, super-res generator (2 frames 144x256)
    for name in ('lres_G', 'lres_D', 'sres_G'):
        a, b = theirs[name], ours[name]
        assert a.shape == b.shape and torch.isfinite(b).all(), name
        err = float((a - b).abs().max()) / max(float(a.abs().max()), 1e-30)
        assert err <= 1e-5, f'{name}: {err:.3e}'
