"""On-device rigid-water check in a process of its own (spawned by tests/test_gpu_settle.py; settle.cu has not run
on hardware yet).  Prints one JSON line; exit code 0 = every check passed."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from molchanica_b200 import workloads as W  # noqa: E402
from molchanica_b200.engine import MdEngine  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from util import trajectory_close  # noqa: E402

D_OH, ANG = 0.9572, np.radians(104.52)
D_HH = float(2 * D_OH * np.sin(ANG / 2))


def main():
    w = dict(W.water_box_c1(), dt=0.001)
    n = len(w["xyzq"])
    triples = np.arange(n, dtype=np.int32).reshape(-1, 3)
    ext = np.asarray(w["box_ext"], np.float64)
    e = MdEngine.from_workload(w)
    e.set_rigid_waters(triples, D_OH, D_HH, 15.999, 1.008)
    n_steps = 60
    e.step(w["dt"], n_steps)
    x = e.positions()
    # velocity stage of RATTLE (known answer): after mc_step every constrained distance is stationary, d/dt |r_ij| = 0
    vv = e.velocities()[:, :3].astype(np.float64).reshape(-1, 3, 3)
    xx = x[:, :3].astype(np.float64).reshape(-1, 3, 3)

    def rdot(i, j):
        r = xx[:, i] - xx[:, j]
        r -= np.rint(r / ext) * ext
        return np.abs((r * (vv[:, i] - vv[:, j])).sum(1)) / np.linalg.norm(r, axis=1)
    bond_speed = float(max(rdot(0, 1).max(), rdot(0, 2).max(), rdot(1, 2).max()))   # A/ps along constrained bonds
    v_typ = float(np.sqrt((vv ** 2).sum(-1)).mean())
    ref = O.md_run(w, n_steps, precision=64, rigid_waters=(triples, D_OH, D_HH))
    ok, worst, sc = trajectory_close(x, ref["xyzq"], w["xyzq"], w["box_ext"])
    m = x[:, :3].astype(np.float64).reshape(-1, 3, 3)

    def d(a, b):
        v = a - b
        return np.linalg.norm(v - np.rint(v / ext) * ext, axis=1)
    geom = np.stack([d(m[:, 0], m[:, 1]), d(m[:, 0], m[:, 2]), d(m[:, 1], m[:, 2])], 1)
    # temperature with 3 degrees of freedom less per rigid molecule, from the oracle's velocities for comparison
    vr = ref["vel"].astype(np.float64)
    ke_ref = 0.5 * ((vr[:, :3] ** 2).sum(1) / vr[:, 3]).sum() / 418.4
    t_ref = 2 * ke_ref / ((3 * n - 3 * len(triples)) * 0.0019872041)
    e.compute_forces()
    res = dict(traj_ok=bool(ok), traj_worst=float(worst), geom_err=float(np.abs(geom - [D_OH, D_OH, D_HH]).max()),
               temperature=float(e.energy()["temperature"]), temperature_ref=float(t_ref),
               bond_speed_over_typical=bond_speed / v_typ)
    e.close()
    # four-site OPC water: SETTLE on (O, H, H) + virtual site M, against the oracle doing the same in fp64
    w4 = W.water_box_opc()  # 216 molecules, L = 18.64: r_c + skin = 9.3 < L / 2
    a, b = w4["vsite_ab"]
    e = MdEngine.from_workload(w4)
    e.set_rigid_waters(w4["rigid_waters"], w4["d_oh"], w4["d_hh"], 15.999, 1.008)
    e.set_virtual_sites(w4["virtual_sites"], a, b)
    e.step(w4["dt"], 40)
    x4 = e.positions()
    ref4 = O.md_run(w4, 40, precision=64, rigid_waters=(w4["rigid_waters"], w4["d_oh"], w4["d_hh"]),
                    virtual_sites=(w4["virtual_sites"], a, b))
    ok4, worst4, _ = trajectory_close(x4, ref4["xyzq"], w4["xyzq"], w4["box_ext"])
    e.compute_forces()
    f4 = e.forces()
    e.close()
    ext4 = np.asarray(w4["box_ext"], np.float64)
    m4 = x4[:, :3].astype(np.float64).reshape(-1, 4, 3)
    mi = lambda v: v - np.rint(v / ext4) * ext4
    msite = np.abs(mi(m4[:, 3] - (m4[:, 0] + a * mi(m4[:, 1] - m4[:, 0]) + b * mi(m4[:, 2] - m4[:, 0])))).max()
    # bonds to hydrogen by SHAKE: C1 water with its O-H bonds as (O, H, H) clusters, the H-H spring kept
    wc = dict(W.water_box_c1(), dt=0.001)
    idx = np.arange(len(wc["xyzq"]), dtype=np.int32).reshape(-1, 3)
    clusters = np.concatenate([idx, np.full((len(idx), 1), -1, np.int32)], 1)
    lengths = np.tile(np.array([[D_OH, D_OH, 1.0]], np.float32), (len(idx), 1))
    e = MdEngine.from_workload(wc)
    e.set_bonded(wc["bonds"], wc["bond_kr0"])
    e.set_hbond_constraints(clusters, lengths)
    e.step(wc["dt"], 40)
    xc = e.positions()
    vc = e.velocities()[:, :3].astype(np.float64).reshape(-1, 3, 3)
    e.close()
    refc = O.md_run(wc, 40, precision=64, with_bonds=True, hbond_constraints=(clusters, lengths))
    okc, worstc, _ = trajectory_close(xc, refc["xyzq"], wc["xyzq"], wc["box_ext"])
    mc = xc[:, :3].astype(np.float64).reshape(-1, 3, 3)
    extc = np.asarray(wc["box_ext"], np.float64)
    dc = lambda a, b: np.linalg.norm((a - b) - np.rint((a - b) / extc) * extc, axis=1)
    rc_ = lambda i, j: (mc[:, i] - mc[:, j]) - np.rint((mc[:, i] - mc[:, j]) / extc) * extc
    shake_bond_speed = max(float(np.abs((rc_(0, k) * (vc[:, 0] - vc[:, k])).sum(1) / np.linalg.norm(rc_(0, k), axis=1)).max()) for k in (1, 2))
    res.update(shake_bond_speed_over_typical=shake_bond_speed / float(np.sqrt((vc ** 2).sum(-1)).mean()))
    res.update(shake_traj_ok=bool(okc), shake_traj_worst=float(worstc),
               shake_len_err=float(max(np.abs(dc(mc[:, 0], mc[:, 1]) - D_OH).max(), np.abs(dc(mc[:, 0], mc[:, 2]) - D_OH).max())))
    res.update(opc_traj_ok=bool(ok4), opc_traj_worst=float(worst4), opc_msite_err=float(msite),
               opc_m_force=float(np.abs(f4[3::4, :3]).max()))
    good = (res["traj_ok"] and res["geom_err"] < 2e-5 and abs(res["temperature"] - res["temperature_ref"]) < 0.01 * res["temperature_ref"] and res["opc_traj_ok"] and
            res["bond_speed_over_typical"] < 2e-5 and res["shake_bond_speed_over_typical"] < 2e-5 and
            res["opc_msite_err"] < 5e-6 and res["opc_m_force"] == 0.0 and res["shake_traj_ok"] and res["shake_len_err"] < 2e-5)
    print(json.dumps(res))
    return 0 if good else 1


if __name__ == "__main__":
    sys.exit(main())
