"""Generates tests/golden/*.npz.

ref_pairs.npz  -- outputs of the REFERENCE's own pair arithmetic (src/cuda/util.cu, cuda.cu),
                  host-compiled unmodified into oracle/_ref/libref_cuda.so by oracle/Makefile.
                  Needs /root/reference (this container only).
md_small.npz   -- inputs + oracle outputs (neighbour CSR, fp64 forces, energies, a short
                  trajectory) for small seeded systems; the CUDA path is checked against these
                  on the GPU box, where /root/reference does not exist.
Run:  python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from molchanica_b200 import workloads as W  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def ref_pairs():
    O.build()
    ref = O.ref_lib()
    assert ref is not None, "oracle/_ref/libref_cuda.so missing: /root/reference is required"
    rng = np.random.default_rng(20260925)
    n = 256
    tgt = rng.uniform(-6, 6, (n, 3)).astype(np.float32)
    dirs = rng.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    src = (tgt + dirs * rng.uniform(0.9, 12.0, (n, 1))).astype(np.float32)
    sigma = rng.uniform(1.0, 3.6, n).astype(np.float32)
    eps = rng.uniform(0.0, 0.3, n).astype(np.float32)
    qs = (rng.normal(0, 0.4, n) * W.COULOMB_SCALE).astype(np.float32)
    qt = (rng.normal(0, 0.4, n) * W.COULOMB_SCALE).astype(np.float32)
    lj = np.zeros((n, 4), np.float32)
    cf = np.zeros((n, 3), np.float32)
    for k in range(n):
        ref.ref_lj_force(tgt[k].ctypes.data, src[k].ctypes.data, float(sigma[k]), float(eps[k]), lj[k].ctypes.data)
        ref.ref_coulomb_force(src[k].ctypes.data, tgt[k].ctypes.data, float(qs[k]), float(qt[k]), cf[k].ctypes.data)
    mi_ext = rng.uniform(10, 80, (n, 3)).astype(np.float32)
    mi_dv = (rng.uniform(-1.6, 1.6, (n, 3)) * mi_ext).astype(np.float32)
    mi_dv[:8] = (np.array([0.5, -0.5, 1.5]) * mi_ext[:8]).astype(np.float32)  # ties
    mi = np.zeros((n, 3), np.float32)
    for k in range(n):
        ref.ref_min_image(mi_ext[k].ctypes.data, mi_dv[k].ctypes.data, mi[k].ctypes.data)
    # the all-pairs kernels, run as the reference runs them (float3 AoS, size_t counts)
    m = 48
    pos = rng.uniform(0, 14, (m, 3)).astype(np.float32)
    pos += (rng.uniform(-0.2, 0.2, (m, 3))).astype(np.float32)
    # keep atoms >= 2.2 A apart so LJ stays finite-sized
    keep = [0]
    for i in range(1, m):
        if np.min(np.linalg.norm(pos[keep] - pos[i], axis=1)) >= 2.2:
            keep.append(i)
    pos = np.ascontiguousarray(pos[keep])
    m = len(pos)
    q = (rng.normal(0, 0.3, m) * W.COULOMB_SCALE).astype(np.float32)
    sig, ep = np.float32(3.2), np.float32(0.15)
    out_lj = np.zeros((m, 3), np.float32)
    out_q = np.zeros((m, 3), np.float32)
    for i in range(m):
        # lj_force_kernel (cuda.cu:73-102) has no self-pair guard: run it per target with that
        # target deleted from the sources (what a caller with disjoint sets would pass)
        srcs = np.ascontiguousarray(np.delete(pos, i, axis=0))
        o = np.zeros((1, 3), np.float32)
        ref.lj_force_kernel(o.ctypes.data_as(C.c_void_p), srcs.ctypes.data_as(C.c_void_p),
                            pos[i:i + 1].ctypes.data_as(C.c_void_p),
                            np.full(m - 1, sig, np.float32).ctypes.data_as(C.c_void_p),
                            np.full(m - 1, ep, np.float32).ctypes.data_as(C.c_void_p), C.c_size_t(m - 1), C.c_size_t(1))
        out_lj[i] = o[0]
        # coulomb_force_kernel (cuda.cu:10-37) indexes ONE charges array with both i_src and i_tgt,
        # so it cannot express two disjoint sets; use the device helper it calls (util.cu:54-63)
        # and accumulate in fp32 in the kernel's order (cuda.cu:28-35)
        qs_i = np.delete(q, i)
        acc = np.zeros(3, np.float32)
        o3 = np.zeros(3, np.float32)
        for j in range(m - 1):
            ref.ref_coulomb_force(srcs[j].ctypes.data, pos[i].ctypes.data, float(qs_i[j]), float(q[i]), o3.ctypes.data)
            acc = (acc + o3).astype(np.float32)
        out_q[i] = acc
    np.savez_compressed(os.path.join(HERE, "ref_pairs.npz"), tgt=tgt, src=src, sigma=sigma, eps=eps, q_src=qs, q_tgt=qt,
                        ref_lj=lj, ref_coulomb=cf, mi_ext=mi_ext, mi_dv=mi_dv, ref_min_image=mi, ap_pos=pos, ap_q=q,
                        ap_sigma=sig, ap_eps=ep, ap_ref_lj=out_lj, ap_ref_coulomb=out_q)
    print("ref_pairs.npz:", n, "pairs,", m, "all-pairs atoms")


def md_small():
    out = {}
    cases = {"lj512": W.lj_fluid(m=8), "water648": W.water_box_c1(), "glob300": W.globule(300, seed=212, name="glob300")}
    for name, w in cases.items():
        start, idx = O.neighbors(w, brute=True)
        f64, sa, en = O.forces(w, (start, idx), precision=64)
        traj = O.md_run(w, 10, precision=32)
        for k in ("xyzq", "vel", "type", "ljtab", "box_lo", "box_ext", "excl_start", "excl_idx", "pairs14"):
            out[f"{name}.{k}"] = np.asarray(w[k])
        out[f"{name}.scalars"] = np.array([w["periodic"], w["rc_lj"], w["rc_q"], w["skin"], w["coul_mode"], w["alpha"],
                                           w["scale14_lj"], w["scale14_q"], w["dt"]], np.float64)
        out[f"{name}.nbr_start"] = start
        out[f"{name}.nbr_idx"] = idx
        out[f"{name}.f64"] = f64
        out[f"{name}.sumabs"] = sa
        out[f"{name}.energy"] = en
        out[f"{name}.x10"] = traj["xyzq"]
        out[f"{name}.v10"] = traj["vel"]
        print(name, len(w["xyzq"]), "atoms,", len(idx), "list entries")
    d = W.docking_c5(n_rec=400, n_lig=12, n_poses=96, seeds=(515, 516, 517))
    for k, v in d.items():
        if k != "name":
            out[f"dock.{k}"] = np.asarray(v)
    out["dock.scores64"], out["dock.abs64"] = O.dock_score(d, precision=64, with_abs=True)
    np.savez_compressed(os.path.join(HERE, "md_small.npz"), **out)


if __name__ == "__main__":
    ref_pairs()
    md_small()
