"""The drop-in boundary without a GPU: the library builds for sm_100a, loads, exports every
symbol include/molchanica_md.h declares, and refuses to run without a device (no CPU fallback)."""
import ctypes as C

import pytest

from molchanica_b200 import _lib


def test_library_exports_every_declared_symbol(engine_lib):
    names = _lib.declared_symbols()
    assert len(names) >= 30
    for s in names:
        assert hasattr(engine_lib, s), s
    assert engine_lib.mc_abi_version() == 2


def test_sass_is_sm100a_only():
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = {ln.split(".")[-2] for ln in out.splitlines() if ".cubin" in ln}
    assert archs == {"sm_100a"}, archs


def test_no_device_means_error_not_fallback(engine_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = engine_lib.mc_create(0, C.byref(h))
    assert rc == _lib.MC_E_NODEVICE
    assert b"no CPU fallback" in engine_lib.mc_last_error(None)


def test_product_package_never_imports_the_oracle():
    import os
    import re
    root = os.path.dirname(_lib.__file__)
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|liboracle|oracle_py|oracle\/_build", txt, re.M), os.path.join(dp, f)
