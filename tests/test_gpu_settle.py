"""Rigid three-site water on the device against the oracle's fp64 SHAKE.  Needs a B200.

STATUS: settle.cu was written after round 1's GPU budget was spent.  Its arithmetic is verified on the host
(tests/test_settle_cpu.py compiles the same settle_terms.h and checks it against a converged fp64 SHAKE); the
kernel plumbing has not run on hardware yet, so the check runs in a process of its own and is allowed to fail
without turning the suite red (xfail, non-strict) until a GPU run has confirmed it."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="settle.cu not yet run on hardware (round-1 GPU budget spent)")
def test_rigid_water_on_device_follows_the_oracle():
    r = subprocess.run([sys.executable, os.path.join(HERE, "settle_gpu_worker.py")], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
