"""SPME reciprocal space on the device against the fp64 restatement.  Needs a B200.

STATUS: pme.cu was written after round 1's GPU budget was spent.  Its arithmetic is verified on the host
(tests/test_pme_cpu.py: the same pme_terms.h against the fp64 SPME, which converges to an exact Ewald sum that
reproduces the Madelung constant); the kernels and the cuFFT plumbing have not run on hardware yet, so the check
runs in a process of its own and is allowed to fail without turning the suite red (xfail, non-strict)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="pme.cu not yet run on hardware (round-1 GPU budget spent)")
def test_spme_on_device_matches_the_restatement():
    r = subprocess.run([sys.executable, os.path.join(HERE, "pme_gpu_worker.py")], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
