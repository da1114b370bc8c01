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


def test_rust_binding_is_in_step_with_the_header():
    """bindings/rust/src/lib.rs is generated from include/molchanica_md.h (tools/gen_rust_binding.py): it must be the
    generator's current output, declare every function of the header, and lay out the two structs like the ctypes
    mirror the GPU tests drive (same field names, same order)."""
    import importlib.util
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_rust_binding", os.path.join(root, "tools", "gen_rust_binding.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text = open(os.path.join(root, "bindings", "rust", "src", "lib.rs")).read()
    assert text == gen.generate(), "run python tools/gen_rust_binding.py"
    fns = set(re.findall(r"pub fn (mc_[a-z0-9_]+)\(", text))
    assert fns == set(_lib.declared_symbols())
    for rust, py in (("McEnergy", _lib.McEnergy), ("McStats", _lib.McStats)):
        body = re.search(r"pub struct " + rust + r" \{(.*?)\n\}", text, re.S).group(1)
        assert re.findall(r"pub (\w+):", body) == [n for n, _ in py._fields_]
