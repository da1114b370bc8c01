"""The kernel SOURCES of bonded.cu, settle.cu, thermostat.cu and pme.cu, compiled unchanged for the host through
tests/cpp/shim/cuda_runtime.h and run one thread at a time (tests/cpp/kernels_host.cpp), against the fp64 oracles.
These device components were written after round 1's GPU budget was spent; this is how their addressing -- slot
maps in a shuffled cell order, minimum images across the box edge, grid indices, atomics, energy reductions --
was exercised without a GPU.  (Races between threads and the cuFFT calls are outside what a serial run can show.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from molchanica_b200 import workloads as W
from oracle import pme_oracle as P

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def K():
    out = os.path.join(HERE, "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libkernels_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas",
                        "-I", os.path.join(HERE, "cpp", "shim"), "-o", so, os.path.join(HERE, "cpp", "kernels_host.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return C.CDLL(so)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _shuffle(n, seed):
    """A cell order that is not the caller's: slot_of_orig and its inverse."""
    rng = np.random.default_rng(seed)
    slot_of_orig = rng.permutation(n).astype(np.int32)
    orig = np.empty(n, np.int32)
    orig[slot_of_orig] = np.arange(n, dtype=np.int32)
    return slot_of_orig, orig


def _pad4(a, dt):
    out = np.zeros((len(a), 4), dt)
    out[:, :a.shape[1]] = a
    return out


@pytest.mark.parametrize("periodic", [False, True])
def test_bonded_kernel(periodic, K, oracle):
    w = W.bonded_globule(300, seed=41)
    n = len(w["xyzq"])
    if periodic:   # put the globule into a periodic box and wrap every atom on its own: bonds straddle the edge
        L = 30.0
        w = dict(w, periodic=True, box_ext=np.full(3, L, np.float32))
        w["xyzq"] = w["xyzq"].copy()
        w["xyzq"][:, :3] = np.mod(w["xyzq"][:, :3] - w["xyzq"][:, :3].min(0) + 11.0, L)
    so, orig = _shuffle(n, 1)
    x_sorted = np.ascontiguousarray(w["xyzq"][orig])
    force = np.zeros((n, 4), np.float32)
    force[:, :3] = 0.25                                      # the kernel ADDS to what the pair kernel left
    e3 = np.zeros(4, np.float64)   # bond, angle, dihedral energies + virial
    ext = np.ascontiguousarray(w["box_ext"], np.float32)
    K.host_bonded(len(w["bonds"]), _p(np.ascontiguousarray(w["bonds"], np.int32)), _p(np.ascontiguousarray(w["bond_kr0"], np.float32)),
                  len(w["angles"]), _p(_pad4(w["angles"], np.int32)), _p(np.ascontiguousarray(w["angle_kt0"], np.float32)),
                  len(w["dihedrals"]), _p(np.ascontiguousarray(w["dihedrals"], np.int32)), _p(_pad4(w["dihedral_prm"], np.float32)),
                  _p(so), _p(x_sorted), _p(ext), int(periodic), _p(force), _p(e3))
    f64, e64 = oracle.bonded(w)
    got = force[so, :3].astype(np.float64) - 0.25             # back in the caller's order
    assert np.abs(got - f64).max() < 3e-5 * np.abs(f64).max()
    assert np.allclose(e3[:3], e64, rtol=3e-5)
    # e3[3], the bonded virial sum_a (r_a - r_ref) . f_a, against -dU/d(lambda) of the oracle's energy under a uniform
    # scaling of positions and box (only the bonds contribute: angles and dihedrals are scale invariant)

    def u(lam):
        x = np.array(w["xyzq"], np.float64)
        x[:, :3] *= lam
        ws = dict(w, xyzq=x.astype(np.float32), box_ext=np.asarray(w["box_ext"], np.float64) * lam)
        return float(np.sum(oracle.bonded(ws)[1]))
    h = 2e-3
    w_fd = -(u(1 + h) - u(1 - h)) / (2 * h)
    assert abs(e3[3] - w_fd) < 2e-3 * abs(w_fd), (e3[3], w_fd)
    assert np.all(force[:, 3] == 0)


def test_settle_and_virtual_site_kernels(K, oracle):
    w = W.water_box_opc(m=4, L=12.5)
    n = len(w["xyzq"])
    ext = np.ascontiguousarray(w["box_ext"], np.float32)
    a, b = w["vsite_ab"]
    dt = np.float32(0.002)
    so, orig = _shuffle(n, 2)
    x0 = w["xyzq"].copy()
    x0[:, :3] = np.mod(x0[:, :3], ext)                      # atoms wrapped one by one
    v = w["vel"].copy()
    x1 = x0.copy()
    x1[:, :3] += v[:, :3] * dt                               # the unconstrained drift
    xs, vs = np.ascontiguousarray(x1[orig]), np.ascontiguousarray(v[orig])
    waters = _pad4(w["rigid_waters"], np.int32)
    sites = np.ascontiguousarray(w["virtual_sites"], np.int32)
    K.host_settle(len(waters), _p(waters), _p(so), _p(xs), _p(vs), C.c_float(15.999), C.c_float(1.008), C.c_float(w["d_oh"]),
                  C.c_float(w["d_hh"]), _p(ext), 1, C.c_float(dt))
    K.host_vsite_construct(len(sites), _p(sites), _p(so), _p(xs), C.c_float(a), C.c_float(b), _p(ext), 1)
    # reference: the oracle's fp64 SHAKE + construction on the same step
    L = oracle.lib()
    xr, vr = x1.copy(), v.copy()
    wt = np.ascontiguousarray(w["rigid_waters"], np.int32)
    L.orc_set_rigid_waters(C.c_int(len(wt)), _p(wt), C.c_float(w["d_oh"]), C.c_float(w["d_hh"]))
    L.orc_set_virtual_sites(C.c_int(len(sites)), _p(sites), C.c_float(a), C.c_float(b))
    try:
        L.orc_shake_waters(_p(x0), _p(xr), _p(vr), _p(ext), C.c_int(1), C.c_float(dt))
        L.orc_vsite_construct(_p(xr), _p(ext), C.c_int(1))
        got_x, got_v = xs[so], vs[so]
        assert np.abs(got_x[:, :3] - xr[:, :3]).max() < 5e-6
        assert np.abs(got_v[:, :3] - vr[:, :3]).max() < 5e-6 / dt
        assert np.array_equal(got_x[:, 3], x0[:, 3]) and np.array_equal(got_v[:, 3], v[:, 3])    # charge and 1/m untouched
        # force redistribution
        rng = np.random.default_rng(3)
        f = rng.normal(0, 8, (n, 4)).astype(np.float32)
        fs = np.ascontiguousarray(f[orig])
        K.host_vsite_spread(len(sites), _p(sites), _p(so), _p(fs), C.c_float(a), C.c_float(b))
        fr = f.copy()
        L.orc_vsite_spread(_p(fr))
        assert np.abs(fs[so][:, :3] - fr[:, :3]).max() < 1e-5 and np.array_equal(fs[so][:, 3], f[:, 3])
    finally:
        L.orc_set_rigid_waters(C.c_int(0), None, C.c_float(0), C.c_float(0))
        L.orc_set_virtual_sites(C.c_int(0), None, C.c_float(0), C.c_float(0))


def test_shake_h_kernel(K, oracle):
    """Bonds to hydrogen in a periodic box, shuffled slots: the kernel (old positions from x' - v dt) against the oracle's
    fp64 SHAKE that is given the true old positions."""
    w = W.water_box_c1()
    n = len(w["xyzq"])
    ext = np.ascontiguousarray(w["box_ext"], np.float32)
    dt = np.float32(0.001)
    idx = np.arange(n, dtype=np.int32).reshape(-1, 3)
    clusters = np.ascontiguousarray(np.concatenate([idx, np.full((len(idx), 1), -1, np.int32)], 1))
    clusters[::2, 2] = -1                                     # every other cluster constrains one bond only
    lengths = np.tile(np.array([[0.9572, 0.9572, 1.0]], np.float32), (len(idx), 1))
    so, orig = _shuffle(n, 5)
    x0 = w["xyzq"].copy()
    x0[:, :3] = np.mod(x0[:, :3], ext)
    v = w["vel"].copy()
    x1 = x0.copy()
    x1[:, :3] += v[:, :3] * dt
    xs, vs = np.ascontiguousarray(x1[orig]), np.ascontiguousarray(v[orig])
    bad = K.host_shake_h(len(clusters), _p(clusters), _p(lengths), _p(so), _p(xs), _p(vs), _p(ext), 1, C.c_float(dt), C.c_float(1e-6))
    assert bad == 0
    L = oracle.lib()
    xr, vr = x1.copy(), v.copy()
    L.orc_set_hbond_constraints(C.c_int(len(clusters)), _p(clusters), _p(lengths))
    try:
        L.orc_shake_h(_p(x0), _p(xr), _p(vr), _p(ext), C.c_int(1), C.c_float(dt))
    finally:
        L.orc_set_hbond_constraints(C.c_int(0), None, None)
    assert np.abs(xs[so][:, :3] - xr[:, :3]).max() < 6e-6
    assert np.abs(vs[so][:, :3] - vr[:, :3]).max() < 6e-6 / dt
    moved = np.abs(xs[so][:, :3] - x1[:, :3]).max(1) > 0
    assert moved[clusters[1::2, 2]].all() and not moved[idx[::2, 2]].any()       # unconstrained hydrogens are left alone


@pytest.mark.parametrize("coul_mode", [1, 2])
def test_between_molecules_energy_kernel(coul_mode, K, oracle):
    """energy_potential_between_mols: protein globule (molecule 0) against every water (molecules 1..): the kernel on the
    engine's kind of list (rows in a shuffled slot order) against a direct fp64 sum over the oracle's neighbour list."""
    w = W.solvated_c3(n_protein=300, n_water=500, L=30.0)
    w = dict(w, coul_mode=coul_mode)
    n = len(w["xyzq"])
    mol = np.zeros(n, np.uint16)
    mol[300:] = 1 + (np.arange(n - 300) // 3).astype(np.uint16)
    start, idx = oracle.neighbors(w)
    so, orig = _shuffle(n, 13)
    xs = np.ascontiguousarray(w["xyzq"][orig])
    ts = np.ascontiguousarray(w["type"][orig].astype(np.uint16))
    counts = (start[1:] - start[:-1])[orig].astype(np.uint32)
    nstart = np.zeros(n, np.uint32)
    nstart[1:] = np.cumsum(counts)[:-1]
    nlist = np.concatenate([so[idx[start[o]:start[o + 1]]] for o in orig]).astype(np.uint32)
    tab = np.ascontiguousarray(w["ljtab"], np.float32)                      # (T, T, 2): sigma, eps
    T = tab.shape[0]
    dev_tab = np.stack([tab[..., 0] ** 2, 24.0 * tab[..., 1]], -1).astype(np.float32)
    ext = np.ascontiguousarray(w["box_ext"], np.float32)
    K.host_between_mols.restype = C.c_double
    got = K.host_between_mols(n, _p(xs), _p(ts), _p(orig), _p(mol), _p(nstart), _p(counts), _p(nlist), _p(dev_tab), T, _p(ext), 1,
                              C.c_float(w["rc_lj"]), C.c_float(w["rc_q"]), 1, coul_mode, C.c_float(0.35))
    from util import between_mols_reference
    want = between_mols_reference(w, mol, start, idx)
    assert want != 0.0 and abs(got - want) < 2e-5 * max(abs(want), 1.0), (got, want)


def test_langevin_kernel(K, oracle):
    rng = np.random.default_rng(8)
    n = 3000
    so, orig = _shuffle(n, 4)
    vel = np.concatenate([rng.normal(0, 3, (n, 3)), 1.0 / rng.uniform(1, 40, (n, 1))], 1).astype(np.float32)   # slot order
    flags = np.zeros(n, np.uint8)
    flags[::40] = 1
    v0 = vel.copy()
    kT, c1 = 0.0019872041 * 310.0, np.exp(-2.0 * 0.002)
    K.host_langevin(n, _p(vel), _p(orig), _p(flags), C.c_float(c1), C.c_float(np.sqrt(1 - c1 * c1)), C.c_float(kT), C.c_uint64(77), C.c_uint64(12))
    L = oracle.lib()
    d = np.zeros(3, np.float64)
    for s in (0, 1, 40, 999, n - 1):
        if flags[s]:
            assert np.array_equal(vel[s], v0[s])
            continue
        L.orc_langevin_normals(C.c_uint64(77), C.c_uint32(int(orig[s])), C.c_uint64(12), _p(d))     # keyed by the ORIGINAL id
        want = c1 * v0[s, :3].astype(np.float64) + np.sqrt(1 - c1 * c1) * np.sqrt(kT * float(v0[s, 3]) * 418.4) * d
        assert np.abs(vel[s, :3] - want).max() < 1e-5 * max(1.0, np.abs(want).max())
    assert np.array_equal(vel[:, 3], v0[:, 3])


def test_dock_filter_kernel(K, engine_lib):
    """The device clash filter against its host twin (mc_dock_filter_poses, same pose_terms.h) and the numpy restatement."""
    from oracle import dock_poses as DP
    d = W.docking_c5(n_rec=1500, n_lig=24, n_poses=64, seeds=(515, 516, 517))
    rec, lig = d["rec"], d["lig"]
    site = rec[:, :3].astype(np.float64).mean(0) + np.array([6.0, 0.0, 0.0])
    near_idx = DP.near_site(rec, None, site, 8.0)
    near, near_c = np.ascontiguousarray(rec[near_idx]), np.ascontiguousarray(d["rec_hphob"][near_idx])
    poses = DP.make_poses(site, 8.0, 4, 60)
    anchor = np.ascontiguousarray(d["lig_anchor"], np.float32)
    rs = np.ascontiguousarray(np.array([[*near[i, :3], 0.0] for i in range(len(near)) if near_c[i] and i % 6 == 0], np.float32))
    ls = np.ascontiguousarray(np.array([[*lig[i, :3], 0.0] for i in range(len(lig)) if d["lig_hphob"][i] and i % 4 == 0], np.float32))
    keep = np.full(len(poses), 7, np.uint8)
    K.host_dock_filter(len(rs), _p(rs), len(ls), _p(ls), _p(anchor), C.c_float(np.float32(1.7) * np.float32(1.1)), len(poses), _p(poses), _p(keep))
    twin = np.zeros(len(poses), np.uint8)
    kept = C.c_int64(0)
    assert engine_lib.mc_dock_filter_poses(len(near), _p(near), _p(near_c), len(lig), _p(lig), _p(d["lig_hphob"]), _p(anchor), 1.7, len(poses),
                                           _p(poses), _p(twin), C.byref(kept)) == 0
    assert np.array_equal(keep, twin) and np.array_equal(keep, DP.filter_poses(near, near_c, lig, d["lig_hphob"], anchor, poses, 1.7))
    assert 0 < int(keep.sum()) < len(keep)


def test_zero_velocities_kernel(K):
    rng = np.random.default_rng(1)
    v = rng.normal(size=(777, 4)).astype(np.float32)
    w0 = v[:, 3].copy()
    K.host_zero_velocities(777, _p(v))
    assert np.all(v[:, :3] == 0) and np.array_equal(v[:, 3], w0)


def test_csvr_kernels(K, oracle):
    rng = np.random.default_rng(12)
    n = 2000
    vel = np.concatenate([rng.normal(0, 3, (n, 3)), 1.0 / rng.uniform(1, 40, (n, 1))], 1).astype(np.float32)
    vel[::30, 3] = 0.0
    vel[::30, :3] = 0.0
    mob = vel[:, 3] > 0
    ke_units = float((0.5 * (vel[mob, :3].astype(np.float64) ** 2).sum(1) / vel[mob, 3]).sum())     # amu A^2/ps^2
    red3 = np.array([0.0, ke_units, float(mob.sum())], np.float64)
    lam = np.zeros(1, np.float32)
    v0 = vel.copy()
    kT, c = 0.0019872041 * 300.0, float(np.exp(-10.0 * 0.002))
    K.host_csvr(n, _p(vel), _p(red3), C.c_double(kT), C.c_double(c), C.c_double(3.0 * 100), C.c_uint64(21), C.c_uint64(6), _p(lam))
    want = oracle.lib().orc_csvr_lambda(ke_units / 418.4, kT, 3.0 * mob.sum() - 300.0, c, 21, 6)
    assert abs(float(lam[0]) - want) < 1e-6 and want != 1.0
    assert np.allclose(vel[:, :3], v0[:, :3] * lam[0], rtol=1e-6, atol=0) and np.array_equal(vel[:, 3], v0[:, 3])


def test_pme_kernels(K):
    rng = np.random.default_rng(6)
    n, Lb, alpha = 400, 22.0, 0.35
    x = rng.uniform(-5, Lb + 5, (n, 3))                       # outside the box too: the wrap is the kernel's job
    q = rng.normal(0, 0.4, n)
    q -= q.mean()
    q[::25] = 0.0                                             # uncharged atoms are skipped
    xyzq = np.concatenate([x, (q * W.COULOMB_SCALE)[:, None]], 1).astype(np.float32)
    lo = np.array([0.5, -1.0, 2.0], np.float32)
    ext = np.array([Lb, Lb + 2, Lb - 1], np.float32)
    Kd = np.array([24, 27, 20], np.int32)                     # odd and even dimensions
    grid = np.zeros(tuple(Kd), np.float32)
    K.host_pme_spread(n, _p(xyzq), _p(Kd), _p(lo), _p(ext), _p(grid))
    ref_grid = P.spread(xyzq, lo, ext, tuple(Kd))
    assert np.abs(grid - ref_grid).max() < 3e-6 * np.abs(ref_grid).max()
    fq = np.fft.rfftn(grid.astype(np.float64)).astype(np.complex64)
    cg = fq.view(np.float32).reshape(fq.shape + (2,)).copy()
    e = np.zeros(4, np.float64)   # energy, -, virial, -
    K.host_pme_convolve(_p(Kd), _p(cg), _p(ext), C.c_float(alpha), _p(e))
    bc = P.influence(tuple(Kd), ext, alpha)
    want = fq.astype(np.complex128) * bc
    got = cg[..., 0].astype(np.float64) + 1j * cg[..., 1]
    assert np.abs(got - want).max() < 1e-5 * np.abs(want).max()
    e_ref, f_ref = P.spme(xyzq, lo, ext, alpha, tuple(Kd))
    assert abs(e[0] - e_ref) < 2e-5 * abs(e_ref)
    phi = np.ascontiguousarray((np.fft.irfftn(got, s=tuple(Kd), axes=(0, 1, 2)) * np.prod(Kd)).astype(np.float32))
    force = np.zeros((n, 4), np.float32)
    force[:, 3] = 7.0
    K.host_pme_gather(n, _p(xyzq), _p(Kd), _p(lo), _p(ext), _p(phi), _p(force))
    assert np.abs(force[:, :3] - f_ref).max() < 2e-5 * np.abs(f_ref).max()
    assert np.all(force[:, 3] == 7.0) and np.all(force[::25, :3] == 0)
    # excluded-pair correction in a shuffled order, periodic box
    w = W.water_box_c1()
    m = len(w["xyzq"])
    so, orig = _shuffle(m, 9)
    xs = np.ascontiguousarray(w["xyzq"][orig])
    fx = np.zeros((m, 4), np.float32)
    ee = np.zeros(4, np.float64)
    es, ei = np.ascontiguousarray(w["excl_start"], np.int32), np.ascontiguousarray(w["excl_idx"], np.int32)
    wext = np.ascontiguousarray(w["box_ext"], np.float32)
    K.host_pme_excl(m, _p(xs), _p(orig), _p(so), _p(es), _p(ei), _p(wext), 1, C.c_float(alpha), _p(fx), _p(ee))
    e_x, f_x = P.excl_correction(w["xyzq"], wext, True, es, ei, alpha)
    assert abs(ee[0] - e_x) < 2e-5 * abs(e_x) and np.abs(fx[so][:, :3] - f_x).max() < 2e-5 * np.abs(f_x).max()
