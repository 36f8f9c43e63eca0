"""Drop-in rule ON CUDA: the reference's own `model/*` (unmodified, staged at oracle/_ref/src) runs forward + backward
(+ an R1-style double backward) on cuda twice -- once over this repository's `torch_utils.ops` (package directory ahead
on PYTHONPATH, the documented drop-in), once over the reference's own ops and its own CUDA plugins built for sm_100a
(oracle #2) -- and the networks' outputs and parameter gradients agree to the north_star tolerances."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'long-video-gan_b200')
SRC = os.path.join(ROOT, 'oracle', '_ref', 'src')

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.isdir(os.path.join(SRC, 'model')), reason='oracle/_ref not staged')]


def _run(pythonpath, out_file, ref_plugins):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(pythonpath), LVG_REF_PLUGINS='1' if ref_plugins else '0')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'dropin_cuda_run.py'), out_file], env=env, cwd=SRC,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:]
    return torch.load(out_file)


@pytest.fixture(scope='module')
def runs(tmp_path_factory):
    d = tmp_path_factory.mktemp('dropin')
    theirs = _run([SRC], str(d / 'ref.pt'), True)
    ours = _run([PKG, SRC], str(d / 'ours.pt'), False)
    return ours, theirs


def test_resolution(runs):
    ours, theirs = runs
    assert ours['where']['bias_act'].startswith(PKG) and ours['where']['plugin'] == 'BiasActPlugin'
    assert theirs['where']['bias_act'].startswith(SRC) and theirs['where']['plugin'] == 'module'


# (key, tolerance relative to max|ref|, L2 tolerance)
CHECKS = [
    ('lres_G', 1e-3, 1e-4), ('lres_G_grad', 1e-2, 1e-3),
    ('lres_D', 1e-3, 1e-4), ('lres_D_r1_gx', 1e-2, 1e-3), ('lres_D_grad', 1e-2, 1e-3),
    # fp16 layers: each op rounds to fp16; 14 layers deep
    ('sres_G', 2e-2, 5e-3), ('sres_G_grad', 5e-2, 2e-2),
    ('sres_D', 2e-2, 1e-2), ('sres_D_grad', 5e-2, 2e-2), ('sres_D_gx', 5e-2, 2e-2),
]


@pytest.mark.parametrize('key,tol_max,tol_l2', CHECKS, ids=[c[0] for c in CHECKS])
def test_reference_networks_on_our_kernels_match_reference_cuda(runs, key, tol_max, tol_l2):
    ours, theirs = runs
    a, b = ours[key].double(), theirs[key].double()
    assert a.shape == b.shape and torch.isfinite(a).all(), key
    assert float(b.abs().max()) > 0, key
    emax = float((a - b).abs().max() / b.abs().max())
    el2 = float((a - b).norm() / b.norm())
    assert emax <= tol_max and el2 <= tol_l2, f'{key}: max {emax:.3e} (tol {tol_max:g}), L2 {el2:.3e} (tol {tol_l2:g})'
