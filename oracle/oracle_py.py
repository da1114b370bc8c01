"""ctypes front-end of the CPU oracle (oracle/md_oracle.c) and of the host-compiled reference
helpers (oracle/_ref/libref_cuda.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def build(quiet=True):
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    r = subprocess.run(["make", "-C", _HERE, "-s"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if not quiet:
        print(r.stdout)


class NbParams(C.Structure):
    _fields_ = [("rc_lj", C.c_float), ("rc_q", C.c_float), ("coul_mode", C.c_int),
                ("alpha", C.c_float), ("lj_on", C.c_int), ("coul_on", C.c_int)]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_neighbors_brute.restype = C.c_int64
        _LIB.orc_neighbors_cell.restype = C.c_int64
        _LIB.orc_min_image.restype = C.c_float
        _LIB.orc_min_image.argtypes = [C.c_float, C.c_float]
        _LIB.orc_dist2.restype = C.c_float
        _LIB.orc_kinetic.restype = C.c_double
        _LIB.orc_md_run.restype = C.c_int
        _LIB.orc_csvr_lambda.restype = C.c_double
        _LIB.orc_csvr_lambda.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint64]
    return _LIB


def ref_lib():
    """The reference's own util.cu/cuda.cu, host-compiled (None when never built)."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libref_cuda.so")
        if not os.path.exists(path):
            return None
        _REF = C.CDLL(path)
        _REF.ref_lj_V.restype = C.c_float
        _REF.ref_lj_V.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        _REF.ref_lj_force.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        _REF.ref_coulomb_force.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        _REF.ref_softening_sq.restype = C.c_float
        _REF.ref_inv_sqrt_pi.restype = C.c_float
    return _REF


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def nb_params(w, lj_on=True, coul_on=True):
    return NbParams(float(w["rc_lj"]), float(w["rc_q"]), int(w["coul_mode"]), float(w.get("alpha", 0.35)),
                    int(lj_on), int(coul_on))


def _excl(w):
    es = w.get("excl_start")
    ei = w.get("excl_idx")
    if es is None or ei is None or len(ei) == 0:
        return None, None
    return np.ascontiguousarray(es, np.int32), np.ascontiguousarray(ei, np.int32)


def neighbors(w, xyzq=None, brute=False):
    """Verlet list (CSR: start int64 n+1, idx int32), rows ascending, radius max(rc)+skin."""
    xyzq = np.ascontiguousarray(w["xyzq"] if xyzq is None else xyzq, np.float32)
    n = len(xyzq)
    ext = np.ascontiguousarray(w["box_ext"], np.float32)
    lo = np.ascontiguousarray(w["box_lo"], np.float32)
    if not w["periodic"]:
        lo = (xyzq[:, :3].min(0) - 0.5).astype(np.float32)
        ext = (xyzq[:, :3].max(0) + 0.5 - lo).astype(np.float32)
    r_list = np.float32(max(w["rc_lj"], w["rc_q"])) + np.float32(w["skin"])
    es, ei = _excl(w)
    start = np.zeros(n + 1, np.int64)
    L = lib()
    if brute:
        def call(idx, cap):
            return L.orc_neighbors_brute(C.c_int(n), _p(xyzq, C.c_float), _p(ext, C.c_float), C.c_int(int(w["periodic"])),
                                         C.c_float(r_list), _p(es, C.c_int32), _p(ei, C.c_int32),
                                         _p(start, C.c_int64), _p(idx, C.c_int32), C.c_int64(cap))
    else:
        def call(idx, cap):
            return L.orc_neighbors_cell(C.c_int(n), _p(xyzq, C.c_float), _p(lo, C.c_float), _p(ext, C.c_float),
                                        C.c_int(int(w["periodic"])), C.c_float(r_list), _p(es, C.c_int32),
                                        _p(ei, C.c_int32), _p(start, C.c_int64), _p(idx, C.c_int32), C.c_int64(cap))
    tot = call(None, 0)
    idx = np.zeros(max(int(tot), 1), np.int32)
    got = call(idx, len(idx))
    assert got == tot, (got, tot)
    return start, idx[:tot]


def forces(w, nbr, xyzq=None, precision=64, lj_on=True, coul_on=True, with_pairs14=True, scale="terms"):
    """(f (n,4) f32 = fx,fy,fz,e_i ; scale (n,) ; energy [E_lj, E_coul] f64).
    scale="terms": per atom sum over pairs of |LJ repulsive| + |LJ attractive| + |Coulomb| term
    magnitudes (the parity scale, tests/util.py); scale="net": sum of the net pair-force magnitudes."""
    xyzq = np.ascontiguousarray(w["xyzq"] if xyzq is None else xyzq, np.float32)
    n = len(xyzq)
    start, idx = nbr
    typ = np.ascontiguousarray(w["type"], np.uint16)
    tab = np.ascontiguousarray(w["ljtab"], np.float32)
    T = tab.shape[0]
    ext = np.ascontiguousarray(w["box_ext"], np.float32)
    p = nb_params(w, lj_on, coul_on)
    f = np.zeros((n, 4), np.float32)
    sa = np.zeros(n, np.float32)
    st = np.zeros(n, np.float32)
    en = np.zeros(2, np.float64)
    idx_c = np.ascontiguousarray(idx, np.int32) if len(idx) else np.zeros(1, np.int32)
    lib().orc_forces(C.c_int(n), _p(xyzq, C.c_float), _p(typ, C.c_uint16), C.c_int(T), _p(tab, C.c_float),
                     _p(ext, C.c_float), C.c_int(int(w["periodic"])), C.byref(p), _p(start, C.c_int64),
                     _p(idx_c, C.c_int32), C.c_int(precision), _p(f, C.c_float), _p(sa, C.c_float), _p(st, C.c_float),
                     _p(en, C.c_double))
    p14 = w.get("pairs14")
    if with_pairs14 and p14 is not None and len(p14):
        p14 = np.ascontiguousarray(p14, np.int32)
        lib().orc_pairs14(C.c_int(len(p14)), _p(p14, C.c_int32), _p(xyzq, C.c_float), _p(typ, C.c_uint16), C.c_int(T),
                          _p(tab, C.c_float), _p(ext, C.c_float), C.c_int(int(w["periodic"])),
                          C.c_float(w["scale14_lj"]), C.c_float(w["scale14_q"]), C.c_int(int(lj_on)),
                          C.c_int(int(coul_on)), _p(f, C.c_float), _p(en, C.c_double), _p(sa, C.c_float),
                          _p(st, C.c_float))
    return f, (st if scale == "terms" else sa), en


def md_run(w, n_steps, precision=32, xyzq=None, vel=None, want_energies=False, ext_force=None,
           with_bonds=False, rigid_waters=None, langevin=None, virtual_sites=None, csvr=None, hbond_constraints=None):
    """n_steps of velocity Verlet on the CPU. Returns dict(xyzq, vel, forces, rebuilds, energies)."""
    xyzq = np.array(w["xyzq"] if xyzq is None else xyzq, np.float32, copy=True)
    vel = np.array(w["vel"] if vel is None else vel, np.float32, copy=True)
    n = len(xyzq)
    typ = np.ascontiguousarray(w["type"], np.uint16)
    tab = np.ascontiguousarray(w["ljtab"], np.float32)
    ext = np.ascontiguousarray(w["box_ext"], np.float32)
    lo = np.ascontiguousarray(w["box_lo"], np.float32)
    if not w["periodic"]:
        # generous static bounding box for the cell grid (coordinates are clamped into it)
        lo = (xyzq[:, :3].min(0) - 8.0).astype(np.float32)
        ext = (xyzq[:, :3].max(0) + 8.0 - lo).astype(np.float32)
    es, ei = _excl(w)
    p14 = w.get("pairs14")
    p14 = np.ascontiguousarray(p14, np.int32) if p14 is not None and len(p14) else None
    bonds = kr0 = None
    if with_bonds and "bonds" in w:
        bonds = np.ascontiguousarray(w["bonds"], np.int32)
        kr0 = np.ascontiguousarray(w["bond_kr0"], np.float32)
    ef = np.ascontiguousarray(ext_force, np.float32) if ext_force is not None else None
    en = np.zeros((n_steps + 1, 4), np.float64) if want_energies else None
    fo = np.zeros((n, 4), np.float32)
    p = nb_params(w)
    wt = None
    if rigid_waters is not None:
        # (triples, d_oh, d_hh): rigid three-site waters, constrained after every drift (orc_shake_waters)
        wt = np.ascontiguousarray(rigid_waters[0], np.int32).reshape(-1, 3)
        lib().orc_set_rigid_waters(C.c_int(len(wt)), _p(wt, C.c_int32), C.c_float(rigid_waters[1]), C.c_float(rigid_waters[2]))
    vs = None
    if virtual_sites is not None:
        # (quads (M, O, H1, H2), a, b): M placed after every drift, its force handed to the parents
        vs = np.ascontiguousarray(virtual_sites[0], np.int32).reshape(-1, 4)
        lib().orc_set_virtual_sites(C.c_int(len(vs)), _p(vs, C.c_int32), C.c_float(virtual_sites[1]), C.c_float(virtual_sites[2]))
    hc = hd = None
    if hbond_constraints is not None:
        # (clusters (m, 4) heavy + up to three hydrogens with -1 for unused, lengths (m, 3)): SHAKE after every drift
        hc = np.ascontiguousarray(hbond_constraints[0], np.int32).reshape(-1, 4)
        hd = np.ascontiguousarray(hbond_constraints[1], np.float32).reshape(-1, 3)
        lib().orc_set_hbond_constraints(C.c_int(len(hc)), _p(hc, C.c_int32), _p(hd, C.c_float))
    if csvr is not None:
        # (temperature_K, 1/tau [1/ps], seed, degrees of freedom removed by constraints): one scale factor per step
        lib().orc_set_csvr(C.c_int(1), C.c_float(csvr[0]), C.c_float(csvr[1]), C.c_uint64(int(csvr[2])), C.c_double(csvr[3] if len(csvr) > 3 else 0.0))
    if langevin is not None:
        # (temperature_K, gamma_per_ps, seed): O step after every drift with Philox noise (orc_langevin_step)
        lib().orc_set_langevin(C.c_int(1), C.c_float(langevin[0]), C.c_float(langevin[1]), C.c_uint64(int(langevin[2])))
    try:
        rb = _md_run_call(n, xyzq, vel, typ, tab, lo, ext, w, p, es, ei, p14, bonds, kr0, ef, n_steps, precision, en, fo)
    finally:
        if hc is not None:
            lib().orc_set_hbond_constraints(C.c_int(0), None, None)
        if csvr is not None:
            lib().orc_set_csvr(C.c_int(0), C.c_float(0), C.c_float(0), C.c_uint64(0), C.c_double(0))
        if vs is not None:
            lib().orc_set_virtual_sites(C.c_int(0), None, C.c_float(0), C.c_float(0))
        if langevin is not None:
            lib().orc_set_langevin(C.c_int(0), C.c_float(0), C.c_float(0), C.c_uint64(0))
        if wt is not None:
            lib().orc_set_rigid_waters(C.c_int(0), None, C.c_float(0), C.c_float(0))
    if rb < 0:
        raise RuntimeError("orc_md_run failed")
    return dict(xyzq=xyzq, vel=vel, forces=fo, rebuilds=rb, energies=en)


def _md_run_call(n, xyzq, vel, typ, tab, lo, ext, w, p, es, ei, p14, bonds, kr0, ef, n_steps, precision, en, fo):
    return lib().orc_md_run(C.c_int(n), _p(xyzq, C.c_float), _p(vel, C.c_float), _p(typ, C.c_uint16),
                          C.c_int(tab.shape[0]), _p(tab, C.c_float), _p(lo, C.c_float), _p(ext, C.c_float),
                          C.c_int(int(w["periodic"])), C.byref(p), C.c_float(w["skin"]), _p(es, C.c_int32),
                          _p(ei, C.c_int32), C.c_int(0 if p14 is None else len(p14)), _p(p14, C.c_int32),
                          C.c_float(w["scale14_lj"]), C.c_float(w["scale14_q"]),
                          C.c_int(0 if bonds is None else len(bonds)), _p(bonds, C.c_int32), _p(kr0, C.c_float),
                          _p(ef, C.c_float), C.c_float(w["dt"]), C.c_int(n_steps), C.c_int(precision),
                          _p(en, C.c_double), _p(fo, C.c_float))


def minimize(w, max_iters):
    """The steepest-descent minimiser of mc_minimize_energy in fp64 (quenched MD moves dx = 418.4 F/m tau^2, tau x 1.2 on an
    accepted move, undone and tau x 0.5 on a rejected one, stop after three accepted moves without change).  The list
    is rebuilt at every trial.  Returns dict(xyzq, accepted, e_initial, e_final, energies)."""
    x = np.array(w["xyzq"], np.float32, copy=True)
    inv_m = w["vel"][:, 3].astype(np.float64)
    if w.get("flags") is not None:
        inv_m = np.where(np.asarray(w["flags"]) & 1, 0.0, inv_m)

    def evaluate(xx):
        ww = dict(w, xyzq=xx)
        f, _, en = forces(ww, neighbors(ww), precision=64)
        e = float(en.sum())
        if w.get("pairs14") is not None and len(w["pairs14"]):
            raise NotImplementedError("oracle minimiser: 1-4 pairs not wired")
        if w.get("bonds") is not None and w.get("_min_bonded"):
            fb, eb = bonded(ww)
            f = f.copy()
            f[:, :3] += fb
            e += float(eb.sum())
        return e, f[:, :3].astype(np.float64)
    e_cur, f = evaluate(x)
    e0, tau, accepted, small, hist = e_cur, 0.001, 0, 0, [e_cur]
    for _ in range(max_iters):
        xt = x.copy()
        xt[:, :3] = (x[:, :3].astype(np.float64) + 418.4 * f * inv_m[:, None] * tau * tau).astype(np.float32)
        e_try, f_try = evaluate(xt)
        if e_try <= e_cur:
            small = small + 1 if (e_cur - e_try) <= 1e-9 * max(1.0, abs(e_cur)) else 0
            x, f, e_cur = xt, f_try, e_try
            accepted += 1
            hist.append(e_cur)
            tau = min(tau * 1.2, 0.02)
            if small >= 3:
                break
        else:
            tau *= 0.5
            if tau < 1e-7:
                break
    return dict(xyzq=x, accepted=accepted, e_initial=e0, e_final=e_cur, energies=np.array(hist))


def bonded(w, xyzq=None):
    """fp64 bonded forces (n,3) and energies {bond, angle, dihedral} of the terms a workload carries
    (keys bonds/bond_kr0, angles/angle_kt0, dihedrals/dihedral_prm; missing kinds count as empty)."""
    x = np.ascontiguousarray(w["xyzq"] if xyzq is None else xyzq, np.float32)
    n = len(x)

    def arr(key, dt, width):
        a = w.get(key)
        if a is None or len(a) == 0:
            return None, 0
        a = np.ascontiguousarray(a, dt).reshape(-1, width)
        return a, len(a)
    b, nb = arr("bonds", np.int32, 2)
    bk, _ = arr("bond_kr0", np.float32, 2)
    a, na = arr("angles", np.int32, 3)
    ak, _ = arr("angle_kt0", np.float32, 2)
    d, nd = arr("dihedrals", np.int32, 4)
    dk, _ = arr("dihedral_prm", np.float32, 3)
    f = np.zeros((n, 3), np.float64)
    e3 = np.zeros(3, np.float64)
    ext = np.ascontiguousarray(w["box_ext"], np.float32)
    lib().orc_bonded64(C.c_int(nb), _p(b, C.c_int32), _p(bk, C.c_float), C.c_int(na), _p(a, C.c_int32), _p(ak, C.c_float),
                       C.c_int(nd), _p(d, C.c_int32), _p(dk, C.c_float), _p(x, C.c_float), _p(ext, C.c_float),
                       C.c_int(int(w["periodic"])), _p(f, C.c_double), _p(e3, C.c_double))
    return f, e3


def dock_score(d, precision=64, poses=None, with_abs=False):
    """(P,5) f32: score, vdw, hydrophobic, electrostatic, coulomb_e
    [+ (P,3) sums of term magnitudes: vdw, coulomb force, coulomb energy]."""
    poses = np.ascontiguousarray(d["poses"] if poses is None else poses, np.float32)
    rec = np.ascontiguousarray(d["rec"], np.float32)
    lig = np.ascontiguousarray(d["lig"], np.float32)
    tab = np.ascontiguousarray(d["ljtab"], np.float32)
    out = np.zeros((len(poses), 5), np.float32)
    out_abs = np.zeros((len(poses), 3), np.float32)
    anchor = np.ascontiguousarray(d["lig_anchor"], np.float32)
    lib().orc_dock_score(C.c_int(len(rec)), _p(rec, C.c_float), _p(np.ascontiguousarray(d["rec_type"], np.uint16), C.c_uint16),
                         _p(np.ascontiguousarray(d["rec_hphob"], np.uint8), C.c_uint8), C.c_int(len(lig)),
                         _p(lig, C.c_float), _p(np.ascontiguousarray(d["lig_type"], np.uint16), C.c_uint16),
                         _p(np.ascontiguousarray(d["lig_hphob"], np.uint8), C.c_uint8), _p(anchor, C.c_float),
                         C.c_int(tab.shape[1]), _p(tab, C.c_float), C.c_int(len(poses)), _p(poses, C.c_float),
                         C.c_int(precision), _p(out, C.c_float), _p(out_abs, C.c_float))
    return (out, out_abs) if with_abs else out
