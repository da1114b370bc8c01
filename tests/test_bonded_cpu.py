"""Bonded terms without a GPU.  The arithmetic the GPU kernel runs (molchanica_b200/csrc/
bonded_terms.h) is compiled for the host into a TEST library and must agree with (1) the independent fp64 oracle
(oracle/md_oracle.c orc_bonded64, a different derivation of the same gradients) and, like the oracle itself,
with (2) central finite differences of the energy; plus known answers that depend on neither."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from molchanica_b200 import workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_math():
    out = os.path.join(HERE, "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libbonded_math_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", so,
                        os.path.join(HERE, "cpp", "bonded_math_host.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return C.CDLL(so)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _host_eval(L, xyz, w):
    xyz = np.ascontiguousarray(xyz, np.float32)
    f = np.zeros_like(xyz)
    e3 = np.zeros(3, np.float64)
    g = lambda k, dt: np.ascontiguousarray(w[k], dt)
    L.bonded_host_eval(_p(xyz), C.c_int64(len(w["bonds"])), _p(g("bonds", np.int32)), _p(g("bond_kr0", np.float32)),
                       C.c_int64(len(w["angles"])), _p(g("angles", np.int32)), _p(g("angle_kt0", np.float32)),
                       C.c_int64(len(w["dihedrals"])), _p(g("dihedrals", np.int32)), _p(g("dihedral_prm", np.float32)),
                       _p(f), _p(e3))
    return f, e3


def test_device_arithmetic_matches_independent_fp64_oracle(host_math, oracle):
    w = W.bonded_globule(400)
    assert len(w["bonds"]) > 300 and len(w["angles"]) > 300 and len(w["dihedrals"]) > 300
    f32_, e32 = _host_eval(host_math, w["xyzq"][:, :3], w)
    f64, e64 = oracle.bonded(w)
    assert np.all(e64 > 0)
    assert np.allclose(e32, e64, rtol=2e-5)
    scale = np.abs(f64).max()
    assert np.abs(f32_ - f64).max() < 2e-5 * scale        # fp32 terms vs fp64 terms, relative to the largest force
    assert np.abs(f64.sum(0)).max() < 1e-9 * scale * len(f64)   # Newton's third law per term


@pytest.mark.parametrize("kind", ["bonds", "angles", "dihedrals"])
def test_oracle_forces_are_minus_the_energy_gradient(kind, oracle):
    w = W.bonded_globule(60, seed=77)
    only = dict(w)
    for k in ("bonds", "angles", "dihedrals"):
        if k != kind:
            only[k] = np.zeros((0, {"bonds": 2, "angles": 3, "dihedrals": 4}[k]), np.int32)
    x0 = w["xyzq"].astype(np.float64)
    # positions are handed over as f32: differentiate on an f32-representable grid with a step well above its ulp
    f, _ = oracle.bonded(only)
    h = 2.0 ** -13
    rng = np.random.default_rng(3)
    for i in rng.choice(len(x0), 12, replace=False):
        for a in range(3):
            xp, xm = x0.copy(), x0.copy()
            xp[i, a] = np.float32(x0[i, a]) + h
            xm[i, a] = np.float32(x0[i, a]) - h
            ep = oracle.bonded(only, xyzq=xp.astype(np.float32))[1].sum()
            em = oracle.bonded(only, xyzq=xm.astype(np.float32))[1].sum()
            fd = -(ep - em) / (2 * h)
            assert abs(fd - f[i, a]) < 2e-4 * max(1.0, np.abs(f).max()), (kind, i, a, fd, f[i, a])


def test_known_answers(host_math, oracle):
    # bond: two atoms 1.2 A apart, r0 = 1.0, k = 100 -> E = 100 * 0.04 = 4, |F| = 2 k dr = 40, restoring
    xyz = np.array([[0, 0, 0], [1.2, 0, 0], [0, 5, 0], [0, 5, 1], [9, 9, 9], [9, 9, 9]], np.float32)
    one = dict(bonds=np.array([[0, 1]], np.int32), bond_kr0=np.array([[100.0, 1.0]], np.float32),
               angles=np.zeros((0, 3), np.int32), angle_kt0=np.zeros((0, 2), np.float32),
               dihedrals=np.zeros((0, 4), np.int32), dihedral_prm=np.zeros((0, 3), np.float32))
    f, e = _host_eval(host_math, xyz, one)
    assert abs(e[0] - 4.0) < 1e-5 and np.allclose(f[0], [40, 0, 0], atol=1e-4) and np.allclose(f[1], [-40, 0, 0], atol=1e-4)
    # angle: right angle, theta0 = 100 deg, k = 50 -> E = 50 (10 deg)^2, the force opens the angle
    xyz = np.array([[1, 0, 0], [0, 0, 0], [0, 1, 0]], np.float32)
    th0 = np.radians(100.0)
    ang = dict(one, bonds=np.zeros((0, 2), np.int32), bond_kr0=np.zeros((0, 2), np.float32),
               angles=np.array([[0, 1, 2]], np.int32), angle_kt0=np.array([[50.0, th0]], np.float32))
    f, e = _host_eval(host_math, xyz, ang)
    assert abs(e[1] - 50.0 * np.radians(10.0) ** 2) < 1e-4
    assert f[0, 1] < 0 and f[2, 0] < 0 and abs(f[0, 0]) < 1e-4 and np.abs(f.sum(0)).max() < 1e-4
    assert abs(abs(f[0, 1]) - 2 * 50.0 * np.radians(10.0)) < 1e-3       # |F| = |dE/dtheta| / |a| with |a| = 1
    # dihedral: trans butane-like chain (phi = 180 deg): n = 3, phase 0 -> E = pk (1 + cos 540) = 0, zero force;
    # cis (phi = 0): E = 2 pk; a 60 degree twist with n = 1: E = pk (1 + cos 60) = 1.5 pk
    def chain(phi_deg):
        p = np.radians(phi_deg)
        return np.array([[1, 1, 0], [1, 0, 0], [0, 0, 0], [0, np.cos(p), np.sin(p)]], np.float32)
    dih = dict(ang, angles=np.zeros((0, 3), np.int32), angle_kt0=np.zeros((0, 2), np.float32),
               dihedrals=np.array([[0, 1, 2, 3]], np.int32), dihedral_prm=np.array([[2.0, 3.0, 0.0]], np.float32))
    f, e = _host_eval(host_math, chain(180.0), dih)
    assert abs(e[2]) < 1e-5 and np.abs(f).max() < 1e-4
    f, e = _host_eval(host_math, chain(0.0), dih)
    assert abs(e[2] - 4.0) < 1e-5
    dih1 = dict(dih, dihedral_prm=np.array([[2.0, 1.0, 0.0]], np.float32))
    for phi in (60.0, -60.0):
        f, e = _host_eval(host_math, chain(phi), dih1)
        assert abs(e[2] - 3.0) < 1e-5
        w = dict(dih1, xyzq=np.concatenate([chain(phi), np.zeros((4, 1), np.float32)], 1), box_ext=np.ones(3, np.float32),
                 periodic=False)
        f64, e64 = oracle.bonded(w)
        assert abs(e64[2] - 3.0) < 1e-6 and np.abs(f - f64).max() < 1e-4
    # the sign convention: a phase of +90 deg distinguishes +60 from -60 (IUPAC: positive = clockwise looking j -> k)
    dihp = dict(dih, dihedral_prm=np.array([[2.0, 1.0, np.pi / 2]], np.float32))
    e_pos = _host_eval(host_math, chain(60.0), dihp)[1][2]
    e_neg = _host_eval(host_math, chain(-60.0), dihp)[1][2]
    wp = dict(dihp, xyzq=np.concatenate([chain(60.0), np.zeros((4, 1), np.float32)], 1), box_ext=np.ones(3, np.float32), periodic=False)
    assert abs(e_pos - oracle.bonded(wp)[1][2]) < 1e-5 and abs(e_pos - e_neg) > 1.0
