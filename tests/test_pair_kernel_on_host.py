"""The hot kernel of the path, on the CPU: pair_force.cu compiled unchanged for the host through the multi-threaded
stand-in tests/cpp/shim_mt/cuda_runtime.h (OS threads + barriers for the block's threads, shuffles and votes) and held
to the same parity bar as on the GPU -- forces within 1e-5 of sum|f_ij| of the fp64 oracle, energy row sums within 1e-5.
A regression test for the lane mapping, the two-gathers-in-flight loop, interior / wrapped rows, the multi-type table
and the decomposed row order that needs no GPU (the GPU suite remains the authority for the compiled device code)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from molchanica_b200 import workloads as W
from util import FORCE_RTOL, force_rel_err

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def K():
    out = os.path.join(HERE, "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libpair_kernel_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-O1", "-std=c++20", "-pthread", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                        "-I", os.path.join(HERE, "cpp", "shim_mt"), "-o", so, os.path.join(HERE, "cpp", "pair_kernel_host.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _interior_flags(w):
    """MC_FLAG_INTERIOR as reorder_kernel sets it: the atom's cell and its 26 neighbours do not touch the box faces."""
    n = len(w["xyzq"])
    if not w["periodic"]:
        return np.zeros(n, np.uint8)
    ext = np.asarray(w["box_ext"], np.float64)
    r_list = max(w["rc_lj"], w["rc_q"]) + w["skin"]
    nc = np.maximum(1, np.floor(ext / (r_list * 1.001 + 1e-3)).astype(int))
    x = np.mod(w["xyzq"][:, :3].astype(np.float64) - np.asarray(w["box_lo"], np.float64), ext)
    k = np.minimum(np.floor(x / ext * nc).astype(int), nc - 1)
    inner = np.all((nc >= 3) & (k >= 1) & (k <= nc - 2), axis=1)
    return np.where(inner, 0x80, 0).astype(np.uint8)


def _run(K, w, variant, oracle, n_interior=0, n_first=0, wrap_positions=True):
    w = dict(w)
    if w["periodic"] and wrap_positions:   # the engine wraps positions into the box at every build
        x = w["xyzq"].copy()
        lo, ext = np.asarray(w["box_lo"], np.float32), np.asarray(w["box_ext"], np.float32)
        x[:, :3] = lo + np.mod(x[:, :3] - lo, ext)
        w["xyzq"] = x
    start, idx = oracle.neighbors(w)
    n = len(w["xyzq"])
    counts = (start[1:] - start[:-1]).astype(np.uint32)
    nstart = start[:-1].astype(np.uint32)
    nlist = np.ascontiguousarray(idx.astype(np.uint32))
    tab = np.asarray(w["ljtab"], np.float32)
    dev_tab = np.ascontiguousarray(np.stack([tab[..., 0] ** 2, 24.0 * tab[..., 1]], -1).astype(np.float32))
    force = np.full((n, 4), 123.0, np.float32)
    flags = _interior_flags(w)
    rc = K.host_pair_force(variant, n, 0, _p(np.ascontiguousarray(w["xyzq"], np.float32)), _p(np.ascontiguousarray(w["type"], np.uint16)),
                           _p(flags), _p(nstart), _p(counts), _p(nlist), _p(dev_tab), tab.shape[0],
                           _p(np.ascontiguousarray(w["box_ext"], np.float32)), int(w["periodic"]), C.c_float(w["rc_lj"]), C.c_float(w["rc_q"]),
                           C.c_float(w.get("alpha", 0.35)), 1, _p(force), n_interior, n_first)
    assert rc == 0
    p14 = w.get("pairs14")
    if p14 is not None and len(p14):
        # the 1-4 rows as mc_set_pairs14 builds them: symmetric CSR in the caller's ids; pairs14_kernel adds to the force
        p14 = np.asarray(p14, np.int64)
        both = np.concatenate([p14, p14[:, ::-1]])
        both = both[np.argsort(both[:, 0], kind="stable")]
        ps = np.zeros(n + 1, np.int32)
        np.add.at(ps, both[:, 0] + 1, 1)
        ps = np.cumsum(ps).astype(np.int32)
        ident = np.arange(n, dtype=np.int32)
        K.host_pairs14(n, _p(np.ascontiguousarray(w["xyzq"], np.float32)), _p(np.ascontiguousarray(w["type"], np.uint16)), _p(ident),
                       _p(ident), _p(ps), _p(np.ascontiguousarray(both[:, 1], np.int32)), _p(dev_tab), tab.shape[0],
                       _p(np.ascontiguousarray(w["box_ext"], np.float32)), int(w["periodic"]), C.c_float(w["scale14_lj"]),
                       C.c_float(w["scale14_q"]), 1, int(w["coul_mode"] != 0), _p(force))
    f64, sumabs, en = oracle.forces(w, (start, idx), precision=64)
    return force, f64, sumabs, en, flags


def test_c4_instantiation_single_type_periodic(K, oracle):
    w = W.lj_fluid(m=12)                                   # 4 x 4 x 4 cells: interior and wrapped rows both occur
    f, f64, sumabs, en, flags = _run(K, w, 0, oracle)
    assert 0 < (flags != 0).sum() < len(flags)
    assert force_rel_err(f, f64, sumabs).max() < FORCE_RTOL
    # 32 lanes per row with energies: same forces, row sums of the pair energies
    f32l, _, _, _, _ = _run(K, W.lj_fluid(m=8), 4, oracle)
    w8 = W.lj_fluid(m=8)
    f64b, sab, enb = oracle.forces(w8, oracle.neighbors(w8), precision=64)
    assert force_rel_err(f32l, f64b, sab).max() < FORCE_RTOL
    assert np.abs(f32l[:, 3] - f64b[:, 3]).max() < 1e-5 * np.abs(f64b[:, 3]).max()      # per-atom pair-energy row sums
    assert abs(0.5 * f32l[:, 3].sum(dtype=np.float64) - enb.sum()) < 1e-5 * abs(enb.sum())


def test_multi_type_vacuum_plain_coulomb_with_energies(K, oracle):
    w = W.globule(300, seed=33)
    f, f64, sumabs, en, _ = _run(K, w, 1, oracle)
    assert force_rel_err(f, f64, sumabs).max() < FORCE_RTOL
    assert abs(0.5 * f[:, 3].sum(dtype=np.float64) - en.sum()) < 1e-5 * max(abs(en.sum()), 0.5 * np.abs(f64[:, 3]).sum())


@pytest.mark.parametrize("variant", [2, 3])
def test_erfc_water_box_four_lanes_and_uniform_loop(variant, K, oracle):
    w = dict(W.water_box_c1(), coul_mode=2)
    f, f64, sumabs, en, _ = _run(K, w, variant, oracle)
    assert force_rel_err(f, f64, sumabs).max() < FORCE_RTOL
    if variant == 2:
        assert abs(0.5 * f[:, 3].sum(dtype=np.float64) - en.sum()) < 1e-5 * max(abs(en.sum()), 0.5 * np.abs(f64[:, 3]).sum())


def test_decomposed_row_order_gives_the_same_forces(K, oracle):
    """Launch rows ordered interior first, then the first and the last layer (the fused-halo launch): same result."""
    w = W.lj_fluid(m=8)
    n = len(w["xyzq"])
    plain, _, _, _, _ = _run(K, w, 0, oracle)
    n_first, n_last = 70, 50
    remapped, _, _, _, _ = _run(K, w, 0, oracle, n_interior=n - n_first - n_last, n_first=n_first)
    assert np.array_equal(plain, remapped)
