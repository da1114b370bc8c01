"""The C++ host mirror of the `dynamics` API surface (include/molchanica_md.hpp) on a real GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_mirror_runs():
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "host_mirror_smoke")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build_cpp_host()
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "molchanica_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and "host mirror ok" in r.stdout, r.stdout + r.stderr
