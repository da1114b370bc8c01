"""Pins the oracle's pair arithmetic to the reference's own src/cuda/util.cu + cuda.cu, host-
compiled unmodified into oracle/_ref/libref_cuda.so (oracle/Makefile), and to the golden vectors
generated from it (tests/golden/make_golden.py -> ref_pairs.npz)."""
import ctypes as C
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_pairs.npz")


def _orc_pairs(oracle, tgt, src, sigma, eps, qs, qt):
    L = oracle.lib()
    n = len(tgt)
    lj = np.zeros((n, 4), np.float32)
    cf = np.zeros((n, 3), np.float32)
    for k in range(n):
        L.orc_pair_lj(tgt[k].ctypes.data_as(C.c_void_p), src[k].ctypes.data_as(C.c_void_p), C.c_float(sigma[k]),
                      C.c_float(eps[k]), lj[k].ctypes.data_as(C.c_void_p))
        L.orc_pair_coulomb(tgt[k].ctypes.data_as(C.c_void_p), src[k].ctypes.data_as(C.c_void_p), C.c_float(qs[k]),
                           C.c_float(qt[k]), cf[k].ctypes.data_as(C.c_void_p))
    return lj, cf


def test_oracle_matches_golden_reference_vectors(oracle):
    g = np.load(GOLD)
    lj, cf = _orc_pairs(oracle, g["tgt"], g["src"], g["sigma"], g["eps"], g["q_src"], g["q_tgt"])
    # same formulas, same fp32 operation order: a few ulp at most
    np.testing.assert_allclose(lj, g["ref_lj"], rtol=4e-6, atol=1e-30)
    np.testing.assert_allclose(cf, g["ref_coulomb"], rtol=4e-6, atol=1e-30)
    mi = np.array([[oracle.lib().orc_min_image(float(d), float(e)) for d, e in zip(dv, ex)]
                   for dv, ex in zip(g["mi_dv"], g["mi_ext"])], np.float32)
    assert np.array_equal(mi, g["ref_min_image"])  # bit-exact


def test_oracle_matches_live_reference_library(oracle):
    ref = oracle.ref_lib()
    if ref is None:
        pytest.skip("oracle/_ref/libref_cuda.so not built (needs /root/reference at build time)")
    g = np.load(GOLD)
    n = len(g["tgt"])
    out4 = np.zeros(4, np.float32)
    out3 = np.zeros(3, np.float32)
    tgt, src = np.ascontiguousarray(g["tgt"]), np.ascontiguousarray(g["src"])  # keep the buffers alive
    for k in range(0, n, 7):
        ref.ref_lj_force(tgt[k:k + 1].ctypes.data, src[k:k + 1].ctypes.data, float(g["sigma"][k]), float(g["eps"][k]),
                         out4.ctypes.data)
        assert np.array_equal(out4, g["ref_lj"][k])
        ref.ref_coulomb_force(src[k:k + 1].ctypes.data, tgt[k:k + 1].ctypes.data, float(g["q_src"][k]),
                              float(g["q_tgt"][k]), out3.ctypes.data)
        assert np.array_equal(out3, g["ref_coulomb"][k])
    assert ref.ref_softening_sq() == np.float32(1e-6)


def test_all_pairs_kernels_of_reference_match_oracle_forces(oracle):
    """lj_force_kernel / coulomb_force_kernel (reference src/cuda/cuda.cu:10-102, all pairs, no
    cutoff) == the oracle's list-based force sum when the list holds every other atom."""
    g = np.load(GOLD)
    pos, q = g["ap_pos"], g["ap_q"]
    n = len(pos)
    w = dict(xyzq=np.concatenate([pos, q[:, None]], 1).astype(np.float32), type=np.zeros(n, np.uint16),
             ljtab=np.array([[[g["ap_sigma"], g["ap_eps"]]]], np.float32), box_lo=np.zeros(3, np.float32),
             box_ext=np.ones(3, np.float32), periodic=False, rc_lj=1e6, rc_q=1e6, skin=0.0, coul_mode=1,
             excl_start=None, excl_idx=None, pairs14=None, scale14_lj=0.5, scale14_q=1 / 1.2)
    start = np.arange(n + 1, dtype=np.int64) * (n - 1)
    idx = np.concatenate([np.delete(np.arange(n, dtype=np.int32), i) for i in range(n)])
    f_lj, sa, _ = oracle.forces(w, (start, idx), precision=32, coul_on=False)
    f_q, sq, _ = oracle.forces(w, (start, idx), precision=32, lj_on=False)
    np.testing.assert_allclose(f_lj[:, :3], g["ap_ref_lj"], rtol=0, atol=2e-5 * float(sa.max()))
    np.testing.assert_allclose(f_q[:, :3], g["ap_ref_coulomb"], rtol=0, atol=2e-5 * float(sq.max()))
