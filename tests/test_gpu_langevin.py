"""Langevin thermostat on the device against the oracle drawing the same Philox noise.  Needs a B200.

STATUS: the O-step kernel was written after round 1's GPU budget was spent.  Its arithmetic is verified on the host
(tests/test_langevin_cpu.py: published Philox known answers, normals shared with the oracle, the kernel body against
the fp64 formula); it has not run on hardware yet, so the check runs in a process of its own and is allowed to fail
without turning the suite red (xfail, non-strict)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="langevin_ou_kernel not yet run on hardware (round-1 GPU budget spent)")
def test_langevin_on_device_follows_the_oracle_with_the_same_noise():
    r = subprocess.run([sys.executable, os.path.join(HERE, "langevin_gpu_worker.py")], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
