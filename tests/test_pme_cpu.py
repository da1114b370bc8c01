"""Ewald / SPME electrostatics without a GPU.
1. The exact Ewald sum of the restatement (oracle/pme_oracle.py) reproduces the Madelung constant of rock salt.
2. Its SPME (the algorithm pme.cu runs, in fp64) converges to the exact reciprocal sum and has F = -dE/dx.
3. The arithmetic the GPU kernels run (molchanica_b200/csrc/pme_terms.h, compiled for the host into a TEST library,
   loops mirroring the kernels, numpy's FFT in between) reproduces that SPME to fp32 accuracy."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from molchanica_b200 import workloads as W
from oracle import pme_oracle as P

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_math():
    out = os.path.join(HERE, "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libpme_math_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", so,
                        os.path.join(HERE, "cpp", "pme_math_host.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    L = C.CDLL(so)
    L.pme_host_excl.restype = C.c_double
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _salt(m=4, a=2.8):
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)
    q = np.where(g.sum(1) % 2 == 0, 1.0, -1.0)
    xyzq = np.concatenate([g * a + 0.37, q[:, None]], 1).astype(np.float32)
    return xyzq, np.full(3, m * a, np.float32)


def test_exact_ewald_gives_the_madelung_constant_of_rock_salt():
    xyzq, ext = _salt()
    a, alpha = 2.8, 0.45
    e_real, _ = P.real_space_brute(xyzq, ext, alpha, rc=0.5 * float(ext[0]))
    e_rec, f_rec = P.ewald_recip_exact(xyzq, ext, alpha, kmax=10)
    e = e_real + e_rec + P.self_energy(xyzq, alpha)
    madelung = -e / (len(xyzq) / 2) * a          # energy per ion pair = -M q^2 / a
    assert abs(madelung - 1.747565) < 2e-4, madelung
    assert np.abs(f_rec).max() < 1e-6            # perfect lattice: no net force from any part


def _charged_box(n=300, L=24.0, seed=5):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, L, (n, 3))
    q = rng.normal(0, 0.4, n)
    q -= q.mean()
    return np.concatenate([x, (q * W.COULOMB_SCALE)[:, None]], 1).astype(np.float32), np.full(3, L, np.float32)


def test_spme_converges_to_the_exact_reciprocal_sum_and_is_a_gradient():
    xyzq, ext = _charged_box()
    alpha = 0.35
    e_ex, f_ex = P.ewald_recip_exact(xyzq, ext, alpha, kmax=9)
    lo = np.zeros(3, np.float32)
    errs = []
    for K in (16, 24, 32):
        e, f = P.spme(xyzq, lo, ext, alpha, (K, K, K))
        errs.append((abs(e - e_ex) / abs(e_ex), np.abs(f - f_ex).max() / np.abs(f_ex).max()))
    assert errs[2][0] < 2e-4 and errs[2][1] < 2e-3          # ~0.75 A spacing
    assert errs[0][1] > errs[1][1] > errs[2][1]              # finer grids are closer
    # forces are minus the gradient of the SPME energy itself (order-4 splines are differentiable)
    K = (24, 24, 24)
    e0, f = P.spme(xyzq, lo, ext, alpha, K)
    h = 1e-4
    for i, a in ((3, 0), (57, 1), (120, 2)):
        xp, xm = xyzq.astype(np.float64), xyzq.astype(np.float64)
        xp[i, a] += h
        xm[i, a] -= h
        fd = -(P.spme(xp, lo, ext, alpha, K)[0] - P.spme(xm, lo, ext, alpha, K)[0]) / (2 * h)
        assert abs(fd - f[i, a]) < 1e-5 * np.abs(f).max() + 1e-7


def test_device_arithmetic_matches_the_fp64_spme(host_math):
    xyzq, ext = _charged_box(n=500, L=26.0, seed=9)
    xyzq[:, :3] += np.float32(3.0)                # positions outside [0, L): the wrap is part of the arithmetic
    lo = np.array([1.5, -2.0, 0.25], np.float32)
    alpha = 0.35
    K = np.array([28, 24, 30], np.int32)
    n = len(xyzq)
    grid = np.zeros(tuple(K), np.float32)
    host_math.pme_host_spread(C.c_int64(n), _p(xyzq), _p(lo), _p(ext), _p(K), _p(grid))
    ref_grid = P.spread(xyzq, lo, ext, tuple(K))
    assert np.abs(grid - ref_grid).max() < 2e-6 * np.abs(ref_grid).max()
    assert abs(grid.sum(dtype=np.float64) - xyzq[:, 3].sum(dtype=np.float64)) < 1e-3   # partition of unity
    bc = np.zeros((K[0], K[1], K[2] // 2 + 1), np.float32)
    host_math.pme_host_influence(_p(K), _p(ext), C.c_float(alpha), _p(bc))
    ref_bc = P.influence(tuple(K), ext, alpha)
    assert np.abs(bc - ref_bc).max() < 5e-6 * ref_bc.max()
    fq = np.fft.rfftn(grid.astype(np.float64))
    phi = (np.fft.irfftn(fq * bc, s=tuple(K), axes=(0, 1, 2)) * np.prod(K)).astype(np.float32)
    f = np.zeros((n, 3), np.float32)
    host_math.pme_host_gather(C.c_int64(n), _p(xyzq), _p(lo), _p(ext), _p(K), _p(np.ascontiguousarray(phi)), _p(f))
    e_ref, f_ref = P.spme(xyzq, lo, ext, alpha, tuple(K))
    assert np.abs(f - f_ref).max() < 1e-5 * np.abs(f_ref).max()
    mult = np.full(K[2] // 2 + 1, 2.0)
    mult[0] = 1.0
    mult[-1] = 1.0
    e = 0.5 * float((bc.astype(np.float64) * np.abs(fq) ** 2 * mult).sum())
    assert abs(e - e_ref) < 1e-5 * abs(e_ref)


def test_excluded_pair_correction(host_math):
    w = W.globule(200, seed=31)
    xyzq = w["xyzq"]
    ext = np.asarray(w["box_ext"], np.float32)
    es, ei = np.ascontiguousarray(w["excl_start"], np.int32), np.ascontiguousarray(w["excl_idx"], np.int32)
    f = np.zeros((len(xyzq), 3), np.float32)
    e = host_math.pme_host_excl(C.c_int64(len(xyzq)), _p(xyzq), _p(ext), 0, _p(es), _p(ei), C.c_float(0.35), _p(f))
    e_ref, f_ref = P.excl_correction(xyzq, ext, False, es, ei, 0.35)
    assert abs(e - e_ref) < 2e-6 * abs(e_ref) and np.abs(f - f_ref).max() < 1e-5 * np.abs(f_ref).max()
    # known answer: one pair at r = 2 A, qq = 1, alpha = 0.35 -> E = -erf(0.7)/2
    from math import erf
    two = np.array([[0, 0, 0, 1.0], [2.0, 0, 0, 1.0]], np.float32)
    es2, ei2 = np.array([0, 1, 2], np.int32), np.array([1, 0], np.int32)
    f2 = np.zeros((2, 3), np.float32)
    e2 = host_math.pme_host_excl(C.c_int64(2), _p(two), _p(ext), 0, _p(es2), _p(ei2), C.c_float(0.35), _p(f2))
    assert abs(e2 + erf(0.7) / 2.0) < 1e-6 and f2[0, 0] == -f2[1, 0]


def test_parameter_suggestion_meets_its_tolerance(engine_lib):
    """mc_pme_suggest (host only): the suggested alpha bounds the real-space tail, the grid is FFT friendly, and SPME with
    these parameters is as accurate as asked for against the exact reciprocal sum."""
    from math import erfc
    xyzq, ext = _charged_box(n=250, L=22.0, seed=3)
    alpha, grid = C.c_float(0), np.zeros(3, np.int32)
    for rc, tol in ((9.0, 1e-4), (12.0, 5e-4), (8.0, 1e-5)):
        assert engine_lib.mc_pme_suggest(rc, tol, _p(ext), C.byref(alpha), _p(grid)) == 0
        a = float(alpha.value)
        assert erfc(a * rc) / rc <= tol * 1.0001 and erfc(0.98 * a * rc) / rc > tol
        for k in grid:
            m = int(k)
            for pr in (2, 3, 5, 7):
                while m % pr == 0:
                    m //= pr
            assert m == 1 and k >= 8
    assert engine_lib.mc_pme_suggest(9.0, 1e-4, _p(ext), C.byref(alpha), _p(grid)) == 0
    e_ex, f_ex = P.ewald_recip_exact(xyzq, ext, float(alpha.value), kmax=10)
    e, f = P.spme(xyzq, np.zeros(3, np.float32), ext, float(alpha.value), tuple(int(k) for k in grid))
    assert np.abs(f - f_ex).max() < 2e-3 * np.abs(f_ex).max() and abs(e - e_ex) < 1e-3 * abs(e_ex)
    assert engine_lib.mc_pme_suggest(-1.0, 1e-4, _p(ext), C.byref(alpha), _p(grid)) != 0
