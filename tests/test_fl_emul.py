"""The fused filtered_lrelu kernel's stage functions compiled as HOST code (csrc/filtered_lrelu_v3.cuh with
FLV3_HOST_EMU) and run thread by thread over a NaN-filled "shared memory" (tools/fl_emul.cu): outputs, sign tensors
(write mode; read mode with offsets and foreign sign-tensor sizes), zero padding bytes and the alignment of every vector
access are checked against the operator's definition (filtered_lrelu.py:121-153 of the reference) without a GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which('nvcc') is None, reason='nvcc not on PATH')
def test_kernel_stages_emulated_on_the_cpu(tmp_path):
    exe = str(tmp_path / 'fl_emul')
    build = subprocess.run(['nvcc', '-O1', '-std=c++17', '-DFLV3_HOST_EMU', '-o', exe, os.path.join(ROOT, 'tools', 'fl_emul.cu')],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert build.returncode == 0, build.stdout[-3000:]
    run = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert run.returncode == 0 and 'all ok' in run.stdout, run.stdout[-3000:]
    # the three configurations of the super-res generator, write / plain / read modes each
    assert run.stdout.count('write:') >= 8 and run.stdout.count('read(') >= 24
