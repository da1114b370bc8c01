"""Rigid three-site water without a GPU.  The SETTLE arithmetic the GPU kernel runs
(molchanica_b200/csrc/settle_terms.h) is compiled for the host into a TEST library and must reproduce a converged
fp64 SHAKE (same constraint equations, independent algorithm), keep the bond lengths and the centre of mass; the
oracle's rigid-water MD (fp64 SHAKE inside orc_md_run) must hold the geometry and conserve energy."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from molchanica_b200 import workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))
D_OH, ANG = 0.9572, np.radians(104.52)
D_HH = 2 * D_OH * np.sin(ANG / 2)
M_O, M_H = 15.999, 1.008


@pytest.fixture(scope="module")
def host_math():
    out = os.path.join(HERE, "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libsettle_math_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", so,
                        os.path.join(HERE, "cpp", "settle_math_host.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return C.CDLL(so)


def _rand_waters(n, rng, sigma):
    base = np.array([[0, 0, 0], [D_OH * np.sin(ANG / 2), D_OH * np.cos(ANG / 2), 0], [-D_OH * np.sin(ANG / 2), D_OH * np.cos(ANG / 2), 0]])
    x0 = np.zeros((n, 3, 3))
    for w in range(n):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        a, b, c, d = q
        R = np.array([[1 - 2 * (c * c + d * d), 2 * (b * c - d * a), 2 * (b * d + c * a)],
                      [2 * (b * c + d * a), 1 - 2 * (b * b + d * d), 2 * (c * d - b * a)],
                      [2 * (b * d - c * a), 2 * (c * d + b * a), 1 - 2 * (b * b + c * c)]])
        x0[w] = base @ R.T + rng.uniform(-30, 30, 3)
    x0 = x0.astype(np.float32).astype(np.float64)
    x1 = (x0 + rng.normal(0, sigma, (n, 3, 3))).astype(np.float32).astype(np.float64)
    return x0, x1


def _shake(x0, x1):
    m = np.array([M_O, M_H, M_H])
    out = x1.copy()
    cons = [(0, 1, D_OH), (0, 2, D_OH), (1, 2, D_HH)]
    for w in range(len(x0)):
        p = out[w]
        for _ in range(1000):
            worst = 0.0
            for i, j, d in cons:
                s, r = p[i] - p[j], x0[w, i] - x0[w, j]
                diff = d * d - s @ s
                worst = max(worst, abs(diff) / (d * d))
                g = diff / (2 * (s @ r) * (1 / m[i] + 1 / m[j]))
                p[i] += g * r / m[i]
                p[j] -= g * r / m[j]
            if worst < 1e-14:
                break
    return out


def _dists(p):
    return np.stack([np.linalg.norm(p[:, 0] - p[:, 1], axis=1), np.linalg.norm(p[:, 0] - p[:, 2], axis=1),
                     np.linalg.norm(p[:, 1] - p[:, 2], axis=1)], 1)


@pytest.mark.parametrize("sigma", [0.002, 0.03, 0.08])
def test_settle_reproduces_converged_shake(sigma, host_math):
    """sigma: rms displacement per coordinate in one step [A]; 0.03 is already far more than a 2 fs step moves a hydrogen."""
    rng = np.random.default_rng(int(sigma * 1e4))
    n = 1500
    x0, x1 = _rand_waters(n, rng, sigma)
    a0 = np.ascontiguousarray(x0.reshape(n, 9), np.float32)
    a1 = np.ascontiguousarray(x1.reshape(n, 9), np.float32)
    out = np.zeros((n, 9), np.float32)
    host_math.settle_host_eval(C.c_int64(n), a0.ctypes.data_as(C.c_void_p), a1.ctypes.data_as(C.c_void_p), C.c_float(M_O),
                               C.c_float(M_H), C.c_float(D_OH), C.c_float(D_HH), out.ctypes.data_as(C.c_void_p))
    o = out.reshape(n, 3, 3).astype(np.float64)
    ref = _shake(x0, x1)
    assert np.abs(_dists(ref) - [D_OH, D_OH, D_HH]).max() < 1e-12
    assert np.abs(_dists(o) - [D_OH, D_OH, D_HH]).max() < 6e-6      # fp32 coordinates of magnitude 30 A: ulp 2e-6
    assert np.abs(o - ref).max() < 6e-6
    m = np.array([M_O, M_H, M_H])[None, :, None]
    assert np.abs((o * m).sum(1) - (x1 * m).sum(1)).max() / (M_O + 2 * M_H) < 4e-6   # centre of mass untouched


def test_kernel_body_recovers_old_positions_and_corrects_velocities(host_math):
    """settle_host_step is the body of settle_kernel line for line: it gets only the drifted positions and the
    velocities (old positions = x' - v dt), works across the periodic boundary, and must land on the SHAKE solution
    with v'' = v + (x'' - x') / dt."""
    rng = np.random.default_rng(11)
    n, dt, L = 1200, 0.002, 25.0
    x0, _ = _rand_waters(n, rng, 0.0)
    x0 = (x0 - x0.min()) % L                      # atoms wrapped one by one: molecules straddle the box edge
    v = rng.normal(0, 6.0, (n, 3, 3))             # A/ps: hot hydrogens, |v dt| ~ 0.012 A
    x0 = x0.astype(np.float32).astype(np.float64)
    v32 = v.astype(np.float32)
    x1 = (x0.astype(np.float32) + v32 * np.float32(dt)).astype(np.float32)
    # reference: unwrap each molecule about its oxygen, SHAKE in fp64, compare displacements
    rel = x0 - x0[:, :1]
    rel -= np.rint(rel / L) * L
    u0 = x0[:, :1] + rel
    u1 = u0 + v32.astype(np.float64) * dt
    ref = _shake(u0, u1)
    xs, vs = x1.reshape(n, 9).copy(), v32.reshape(n, 9).copy()
    ext = np.full(3, L, np.float32)
    host_math.settle_host_step(C.c_int64(n), xs.ctypes.data_as(C.c_void_p), vs.ctypes.data_as(C.c_void_p),
                               ext.ctypes.data_as(C.c_void_p), C.c_float(M_O), C.c_float(M_H), C.c_float(D_OH), C.c_float(D_HH),
                               C.c_float(dt))
    moved = xs.reshape(n, 3, 3).astype(np.float64) - x1.astype(np.float64)
    assert np.abs(moved - (ref - u1)).max() < 8e-6
    dv = vs.reshape(n, 3, 3).astype(np.float64) - v32.astype(np.float64)
    assert np.abs(dv - (ref - u1) / dt).max() < 8e-6 / dt
    # the constrained velocities have no component along the bonds at the half step: d/dt |r_ij|^2 = 0 to first order
    xn = xs.reshape(n, 3, 3).astype(np.float64)
    r01 = xn[:, 0] - xn[:, 1]
    r01 -= np.rint(r01 / L) * L
    assert np.abs(np.linalg.norm(r01, axis=1) - D_OH).max() < 8e-6


def test_oracle_rigid_water_md_holds_geometry_and_energy(oracle):
    w = W.water_box_c1()
    n = len(w["xyzq"])
    triples = np.arange(n, dtype=np.int32).reshape(-1, 3)
    w = dict(w, dt=0.001)
    # the workload's waters are built with the TIP3P geometry
    x = w["xyzq"][:, :3].astype(np.float64).reshape(-1, 3, 3)
    ext = np.asarray(w["box_ext"], np.float64)

    def geom(x):
        d = lambda a, b: np.linalg.norm((a - b) - np.rint((a - b) / ext) * ext, axis=1)
        return np.stack([d(x[:, 0], x[:, 1]), d(x[:, 0], x[:, 2]), d(x[:, 1], x[:, 2])], 1)
    assert np.abs(geom(x) - [D_OH, D_OH, D_HH]).max() < 1e-4
    r = oracle.md_run(w, 120, precision=64, want_energies=True, rigid_waters=(triples, D_OH, D_HH))
    x = r["xyzq"][:, :3].astype(np.float64).reshape(-1, 3, 3)
    assert np.abs(geom(x) - [D_OH, D_OH, D_HH]).max() < 1e-5
    e = r["energies"]
    tot = e[:, 0] + e[:, 1] + e[:, 3]
    # the first constrained step removes the bond-direction components of the Maxwell-Boltzmann velocities; after
    # that the total energy of the rigid-water NVE run stays put (|E_kin| ~ 390 kcal/mol at 300 K)
    assert np.abs(tot[20:] - tot[20]).max() < 0.02 * e[20:, 3].mean(), (tot[20], tot[-1], e[20:, 3].mean())
    # rigid molecules: relative velocity along every bond vanishes at the half step -> at integer steps it is small
    assert r["rebuilds"] >= 1


def test_virtual_site_keeps_total_force_and_torque(host_math):
    rng = np.random.default_rng(4)
    n, a = 500, 0.14773
    x0, _ = _rand_waters(n, rng, 0.0)
    x = np.concatenate([x0, np.zeros((n, 1, 3))], 1).astype(np.float32).reshape(n, 12)
    f = rng.normal(0, 10, (n, 4, 3)).astype(np.float32).reshape(n, 12)
    f_in = f.reshape(n, 4, 3).astype(np.float64).copy()
    host_math.vsite_host_eval(C.c_int64(n), x.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p), C.c_float(a), C.c_float(a))
    xs, fs = x.reshape(n, 4, 3).astype(np.float64), f.reshape(n, 4, 3).astype(np.float64)
    # M sits on the bisector, 2 a cos(theta/2) d_OH from the oxygen
    dm = np.linalg.norm(xs[:, 3] - xs[:, 0], axis=1)
    assert np.abs(dm - 2 * a * D_OH * np.cos(ANG / 2)).max() < 1e-5
    assert np.abs(fs[:, 3]).max() == 0.0
    assert np.abs(fs.sum(1) - f_in.sum(1)).max() < 1e-4                                  # total force
    tq = lambda ff: np.cross(xs - xs[:, :1], ff).sum(1)                                   # torque about the oxygen
    assert np.abs(tq(fs) - tq(f_in)).max() < 2e-4


def test_oracle_four_site_water_md(oracle):
    w = W.water_box_opc(m=5, L=15.6)
    a, b = w["vsite_ab"]
    r = oracle.md_run(w, 100, precision=64, want_energies=True, rigid_waters=(w["rigid_waters"], w["d_oh"], w["d_hh"]),
                      virtual_sites=(w["virtual_sites"], a, b))
    x = r["xyzq"][:, :3].astype(np.float64).reshape(-1, 4, 3)
    ext = np.asarray(w["box_ext"], np.float64)

    def mi(v):
        return v - np.rint(v / ext) * ext
    d1, d2 = mi(x[:, 1] - x[:, 0]), mi(x[:, 2] - x[:, 0])
    assert np.abs(np.linalg.norm(d1, axis=1) - w["d_oh"]).max() < 1e-5 and np.abs(np.linalg.norm(mi(x[:, 1] - x[:, 2]), axis=1) - w["d_hh"]).max() < 1e-5
    assert np.abs(mi(x[:, 3] - (x[:, 0] + a * d1 + b * d2))).max() < 2e-6                # M follows its parents
    assert np.abs(r["vel"][3::4, :3]).max() == 0.0                                        # and is never integrated
    e = r["energies"]
    tot = e[:, 0] + e[:, 1] + e[:, 3]
    assert np.abs(tot[20:] - tot[20]).max() < 0.03 * e[20:, 3].mean(), (tot[20], tot[-1], e[20:, 3].mean())


def test_hydrogen_bond_shake_header_and_oracle(host_math, oracle):
    """Clusters of a heavy atom with 1-3 hydrogens: the device arithmetic (shake_terms.h, Gauss-Seidel in fp32 to 1e-6)
    and the oracle (fp64, opposite sweep order, 1e-13) land on the same constrained positions; lengths hold; the
    centre of mass of every cluster is untouched."""
    rng = np.random.default_rng(21)
    n = 900
    nh = rng.integers(1, 4, n).astype(np.int32)
    m_heavy = rng.choice([12.011, 14.007, 15.999], n)
    inv_m = np.zeros((n, 4), np.float32)
    inv_m[:, 0] = 1.0 / m_heavy
    inv_m[:, 1:] = 1.0 / 1.008
    d = rng.uniform(0.95, 1.12, (n, 3)).astype(np.float32)
    x0 = np.zeros((n, 4, 3))
    x0[:, 0] = rng.uniform(-25, 25, (n, 3))
    for k in range(3):
        u = rng.normal(size=(n, 3))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        x0[:, 1 + k] = x0[:, 0] + u * d[:, k:k + 1]
    x0 = x0.astype(np.float32).astype(np.float64)
    x1 = (x0 + rng.normal(0, 0.03, x0.shape)).astype(np.float32).astype(np.float64)
    out = np.zeros((n, 12), np.float32)
    a0, a1 = np.ascontiguousarray(x0.reshape(n, 12), np.float32), np.ascontiguousarray(x1.reshape(n, 12), np.float32)
    sweeps = host_math.shake_host_eval(C.c_int64(n), nh.ctypes.data_as(C.c_void_p), a0.ctypes.data_as(C.c_void_p),
                                       a1.ctypes.data_as(C.c_void_p), inv_m.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                                       C.c_float(1e-6), out.ctypes.data_as(C.c_void_p))
    assert 1 < sweeps <= 64
    o = out.reshape(n, 4, 3).astype(np.float64)
    # the oracle on the same clusters, laid out as a 4n-atom system
    xo = np.zeros((4 * n, 4), np.float32)
    xn = np.zeros((4 * n, 4), np.float32)
    vel = np.zeros((4 * n, 4), np.float32)
    xo[:, :3], xn[:, :3] = x0.reshape(-1, 3), x1.reshape(-1, 3)
    vel[:, 3] = inv_m.reshape(-1)
    clusters = np.arange(4 * n, dtype=np.int32).reshape(n, 4)
    for k in range(3):
        clusters[nh <= k, 1 + k] = -1
    L = oracle.lib()
    L.orc_set_hbond_constraints(C.c_int(n), clusters.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p))
    try:
        L.orc_shake_h(xo.ctypes.data_as(C.c_void_p), xn.ctypes.data_as(C.c_void_p), vel.ctypes.data_as(C.c_void_p), None, C.c_int(0),
                      C.c_float(0.002))
    finally:
        L.orc_set_hbond_constraints(C.c_int(0), None, None)
    ref = xn[:, :3].astype(np.float64).reshape(n, 4, 3)
    mass = 1.0 / inv_m.astype(np.float64)
    for k in range(3):
        live = nh > k
        assert np.abs(np.linalg.norm(o[live, 1 + k] - o[live, 0], axis=1) - d[live, k]).max() < 6e-6
        assert np.abs(o[live, 1 + k] - ref[live, 1 + k]).max() < 8e-6
        assert np.array_equal(out.reshape(n, 4, 3)[~live, 1 + k], a1.reshape(n, 4, 3)[~live, 1 + k])   # unused slots untouched
        mass[~live, 1 + k] = 0.0
    assert np.abs(o[:, 0] - ref[:, 0]).max() < 8e-6
    com = lambda p: (p * mass[:, :, None]).sum(1) / mass.sum(1)[:, None]
    assert np.abs(com(o) - com(x1)).max() < 5e-6
    # the velocity correction of the oracle is the position change over dt
    assert np.allclose(vel[:, :3].reshape(n, 4, 3)[:, 0], (ref[:, 0] - x1[:, 0]) / 0.002, atol=2e-3)


def test_oracle_md_with_hydrogen_constraints(oracle):
    """C1 water with its two O-H bonds constrained as one (O, H, H) cluster and the H-H spring kept (a flexible angle):
    1 fs steps, the constrained lengths hold and the total energy stays put once the first step has removed the
    bond-direction velocities."""
    w = dict(W.water_box_c1(), dt=0.001)
    n = len(w["xyzq"])
    idx = np.arange(n, dtype=np.int32).reshape(-1, 3)
    clusters = np.concatenate([idx, np.full((len(idx), 1), -1, np.int32)], 1)
    lengths = np.tile(np.array([[D_OH, D_OH, 1.0]], np.float32), (len(idx), 1))
    r = oracle.md_run(w, 120, precision=64, want_energies=True, with_bonds=True, hbond_constraints=(clusters, lengths))
    x = r["xyzq"][:, :3].astype(np.float64).reshape(-1, 3, 3)
    ext = np.asarray(w["box_ext"], np.float64)
    d = lambda a, b: np.linalg.norm((a - b) - np.rint((a - b) / ext) * ext, axis=1)
    assert np.abs(d(x[:, 0], x[:, 1]) - D_OH).max() < 1e-5 and np.abs(d(x[:, 0], x[:, 2]) - D_OH).max() < 1e-5
    assert np.abs(d(x[:, 1], x[:, 2]) - D_HH).max() > 1e-3          # the angle is free
    e = r["energies"]
    tot = e[:, 0] + e[:, 1] + e[:, 2] + e[:, 3]
    assert np.abs(tot[20:] - tot[20]).max() < 0.02 * e[20:, 3].mean(), (tot[20], tot[-1], e[20:, 3].mean())
