"""Bonded terms on the device against the fp64 oracle.  Needs a B200.

STATUS: bonded.cu was written after round 1's GPU budget was spent.  Its arithmetic is verified on the host
(tests/test_bonded_cpu.py compiles the same bonded_terms.h and checks it against the independent fp64 oracle and
finite differences); the kernel plumbing has not run on hardware yet.  The checks therefore run in a process of
their own (a faulting kernel must not poison the CUDA context of the other GPU tests) and are allowed to fail
without turning the suite red (xfail, non-strict) until a GPU run has confirmed them."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="bonded.cu not yet run on hardware (round-1 GPU budget spent)")
def test_bonded_terms_on_device_match_oracle():
    r = subprocess.run([sys.executable, os.path.join(HERE, "bonded_gpu_worker.py")], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
