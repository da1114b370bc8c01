"""Rigid three-site water on the device against the oracle's fp64 SHAKE.  Needs a B200.

The arithmetic is also verified on the host (tests/test_settle_cpu.py compiles the same settle_terms.h and checks it
against a converged fp64 SHAKE).  The worker script is shared with tests/test_library_on_host.py.  Confirmed on hardware
at the end of round 1: a failure here is a regression and turns the suite red."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_rigid_water_on_device_follows_the_oracle():
    r = subprocess.run([sys.executable, os.path.join(HERE, "settle_gpu_worker.py")], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
