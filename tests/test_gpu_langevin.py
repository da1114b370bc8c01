"""Langevin thermostat on the device against the oracle drawing the same Philox noise.  Needs a B200.

The arithmetic is also verified on the host (tests/test_langevin_cpu.py: published Philox known answers, normals shared
with the oracle, the kernel body against the fp64 formula).  The worker script is shared with
tests/test_library_on_host.py.  Confirmed on hardware at the end of round 1: a failure here turns the suite red."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_langevin_on_device_follows_the_oracle_with_the_same_noise():
    r = subprocess.run([sys.executable, os.path.join(HERE, "langevin_gpu_worker.py")], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
