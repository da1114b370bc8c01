"""SPME reciprocal space on the device against the fp64 restatement.  Needs a B200.

The arithmetic is also verified on the host (tests/test_pme_cpu.py: the same pme_terms.h against the fp64 SPME, which
converges to an exact Ewald sum that reproduces the Madelung constant).  The worker script is shared with
tests/test_library_on_host.py.  Kernels and cuFFT plumbing confirmed on hardware at the end of round 1: a failure here
turns the suite red."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_spme_on_device_matches_the_restatement():
    r = subprocess.run([sys.executable, os.path.join(HERE, "pme_gpu_worker.py")], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
