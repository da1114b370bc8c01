"""Bonded terms on the device against the fp64 oracle.  Needs a B200.

The arithmetic is also verified on the host (tests/test_bonded_cpu.py compiles the same bonded_terms.h and checks it
against the independent fp64 oracle and finite differences).  The worker runs in a process of its own (its checks are
shared with tests/test_library_on_host.py, which runs the same script against the host build).  Confirmed on hardware at
the end of round 1: a failure here is a regression and turns the suite red."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_bonded_terms_on_device_match_oracle():
    r = subprocess.run([sys.executable, os.path.join(HERE, "bonded_gpu_worker.py")], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
