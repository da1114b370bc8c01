"""Runs the on-device bonded-term checks in a process of its own (spawned by tests/test_gpu_bonded.py): bonded.cu
has not run on hardware yet, and a faulting kernel must not poison the CUDA context of the other GPU tests.
Prints one JSON line with the measured errors; exit code 0 = every check passed."""
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
from util import FORCE_RTOL, between_mols_reference, trajectory_close  # noqa: E402


def main():
    res = {}
    # 1. forces and energies of bonds + angles + dihedrals + nonbonded against the fp64 oracle
    w = W.bonded_globule(400)
    e = MdEngine.from_workload(w, bonded=True)
    e.compute_forces()
    f = e.forces()
    en = e.energy()
    nb = O.neighbors(w)
    f_nb, scale_nb, _ = O.forces(w, nb, precision=64)
    f_b, e_b = O.bonded(w)
    want = f_nb[:, :3] + f_b
    scale = scale_nb + np.abs(f_b).max(1)
    scale = np.maximum(scale, 1e-3 * scale.max())
    res["force_err"] = float((np.abs(f[:, :3].astype(np.float64) - want).max(1) / scale).max())
    res["energy_rel"] = float(np.abs(np.array([en["energy_bond"], en["energy_angle"], en["energy_dihedral"]]) - e_b).max() / e_b.max())
    res["sum_ok"] = bool(abs(en["energy_potential"] - (en["energy_potential_nonbonded"] + en["energy_potential_bonded"])) < 1e-9)
    e.close()
    e = MdEngine.from_workload(w)
    e.compute_forces()
    res["no_terms_zero"] = bool(e.energy()["energy_potential_bonded"] == 0.0)
    # 2. out-of-range ids are an error
    try:
        e.set_bonded(np.array([[0, len(w["xyzq"])]], np.int32), np.array([[100.0, 1.0]], np.float32))
        res["bad_id_rejected"] = False
    except Exception:
        res["bad_id_rejected"] = True
    e.close()
    # 3. C1: 216 flexible three-site waters, 40 NVE steps against the oracle with the same bonds
    w = W.water_box_c1()
    e = MdEngine.from_workload(w)
    e.set_bonded(w["bonds"], w["bond_kr0"])
    e.step(w["dt"], 40)
    ref = O.md_run(w, 40, precision=64, with_bonds=True)
    ok, worst, sc = trajectory_close(e.positions(), ref["xyzq"], w["xyzq"], w["box_ext"])
    en = e.energy()
    res.update(traj_ok=bool(ok), traj_worst=float(worst), density=float(en["density"]), e_bond=float(en["energy_bond"]))
    e.close()
    # 4. interaction energy between molecules (protein against every water), on demand
    ws = W.solvated_c3(n_protein=300, n_water=500, L=30.0)
    mol = np.zeros(len(ws["xyzq"]), np.uint16)
    mol[300:] = 1 + (np.arange(len(mol) - 300) // 3).astype(np.uint16)
    e = MdEngine.from_workload(ws)
    e.compute_forces()
    e_between = e.energy_between_mols(mol)
    e.close()
    s_, i_ = O.neighbors(ws)
    want_b = between_mols_reference(ws, mol, s_, i_)
    res["between_rel"] = float(abs(e_between - want_b) / max(abs(want_b), 1.0))
    # 5. mc_minimize_energy against the fp64 restatement of the same minimiser: energies go down, velocities survive
    wm = W.lj_fluid(m=8, temp_k=50.0)
    e = MdEngine.from_workload(wm)
    v_before = e.velocities()
    acc, e0, e1 = e.minimize_energy(40)
    v_after = e.velocities()
    x_after = e.positions()
    e.close()
    rm = O.minimize(dict(wm, pairs14=None), 40)
    res.update(min_accepted=acc, min_e0=e0, min_e1=e1, min_ref_e1=float(rm["e_final"]),
               min_vel_kept=bool(np.array_equal(v_before, v_after)), min_moved=float(np.abs(x_after[:, :3] - wm["xyzq"][:, :3]).max()))
    min_ok = (acc >= 10 and e1 < e0 and abs(e0 - rm["e_initial"]) < 1e-4 * abs(e0) and
              abs(e1 - rm["e_final"]) < 0.02 * abs(rm["e_initial"] - rm["e_final"]) + 1e-4 * abs(e1) and res["min_vel_kept"])
    # 6. the clash pre-filter of the docking scan on the device against its host twin
    d = W.docking_c5(n_rec=1500, n_lig=24, n_poses=64, seeds=(515, 516, 517))
    e = MdEngine(0)
    site = d["rec"][:, :3].astype(np.float64).mean(0) + np.array([6.0, 0.0, 0.0])
    poses = e.dock_make_poses(site, 8.0, 4, 60)
    k_gpu = e.dock_filter_poses(d["rec"], d["rec_hphob"], d["lig"], d["lig_hphob"], d["lig_anchor"], poses, gpu=True)
    k_cpu = e.dock_filter_poses(d["rec"], d["rec_hphob"], d["lig"], d["lig_hphob"], d["lig_anchor"], poses, gpu=False)
    e.close()
    res["filter_equal"] = bool(np.array_equal(k_gpu, k_cpu))
    res["filter_kept"] = int(k_gpu.sum())
    good = (res["filter_equal"] and 0 < res["filter_kept"] < len(poses) and min_ok and res["between_rel"] < 2e-5 and res["force_err"] < 2 * FORCE_RTOL and res["energy_rel"] < 2e-5 and res["sum_ok"] and res["no_terms_zero"] and
            res["bad_id_rejected"] and res["traj_ok"] and 0.9 < res["density"] < 1.1 and res["e_bond"] > 0)
    print(json.dumps(res))
    return 0 if good else 1


if __name__ == "__main__":
    sys.exit(main())
