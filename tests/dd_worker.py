"""One rank of a domain-decomposed run (spawned by tests/test_gpu_multi.py, one process per GPU).
usage: dd_worker.py <rank> <world> <id_file> <case> <out_npz> [fused|nccl] [fixed|adaptive|allgather]
DD_EXT=<k>: the steps are taken in calls of k steps, each with its own external-force array (ext_forces_for_call), through
the pipelined upload (option defer_tail = DD_DEFER, default 1): a decomposed rank uploads its 1/N block of the array and the
blocks are all-gathered over NCCL."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from molchanica_b200 import workloads as W  # noqa: E402
from molchanica_b200.engine import MdEngine  # noqa: E402


def case_workload(name, world=2):
    if name == "lj":
        # two cell layers per rank are the minimum: 24^3 atoms give 8 layers, 48^3 give 16
        return W.lj_fluid(m=24 if world <= 4 else 48), 30
    if name == "solv":
        w = W.solvated_c3()
        w["coul_mode"] = 2  # continuous at the cutoff: trajectories are comparable
        return w, 6
    if name in ("solvb", "solvb_small", "solvl", "solvl_small", "solvc", "solvc_small"):
        # bonded terms (and, solvl / solvc, the Langevin / CSVR thermostat) on a decomposed handle: every rank evaluates the terms that touch
        # its owned atoms; compared with the single-handle run of the same library (tests/test_gpu_multi.py)
        return W.solvated_bonded(small=name.endswith("_small")), 8
    if name == "ljx":
        # fuzzing hook (tests/test_library_on_host.py, manual sweeps): DD_M atoms per edge, DD_TEMP K, DD_SKIN A, DD_STEPS
        w = W.lj_fluid(m=int(os.environ.get("DD_M", "24")), temp_k=float(os.environ.get("DD_TEMP", "86.3")))
        w["skin"] = float(os.environ.get("DD_SKIN", w["skin"]))
        return w, int(os.environ.get("DD_STEPS", "30"))
    raise ValueError(name)


def ext_forces_for_call(n, k):
    """external forces of call k: every 7th atom, offset k, seeded -- the same arrays on every rank and in the test"""
    ext = np.zeros((n, 3), np.float32)
    ext[k % 7::7] = np.random.default_rng(1000 + k).normal(0, 3.0, ext[k % 7::7].shape)
    return ext


def main():
    rank, world, id_file, case, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
    halo = sys.argv[6] if len(sys.argv) > 6 else "fused"
    sched = sys.argv[7] if len(sys.argv) > 7 else "fixed"
    w, n_steps = case_workload(case, world)
    e = MdEngine(device=rank)
    uid = np.zeros(128, np.uint8)
    if rank == 0:
        e._chk(e._L.mc_comm_unique_id(uid.ctypes.data_as(C.c_void_p)))
        with open(id_file + ".tmp", "wb") as f:
            f.write(uid.tobytes())
        os.replace(id_file + ".tmp", id_file)
    else:
        t0 = time.time()
        while not os.path.exists(id_file):
            if time.time() - t0 > 120:
                raise SystemExit("timed out waiting for the NCCL id")
            time.sleep(0.05)
        uid = np.frombuffer(open(id_file, "rb").read(), np.uint8).copy()
    e._chk(e._L.mc_comm_init(e._h, uid.ctypes.data_as(C.c_void_p), rank, world))
    lo = np.asarray(w["box_lo"], np.float32)
    e.set_box(lo, lo + np.asarray(w["box_ext"], np.float32), True)
    e.set_cutoffs(w["rc_lj"], w["rc_q"], w["skin"], w["coul_mode"], w.get("alpha", 0.35))
    e.set_lj_table(w["ljtab"])
    e.set_atoms(w["xyzq"], w["type"], w["vel"])
    e.set_exclusions(w.get("excl_start"), w.get("excl_idx"))
    e.set_pairs14(w.get("pairs14"), w.get("scale14_lj", 0.5), w.get("scale14_q", 1 / 1.2))
    bonded = case.startswith(("solvb", "solvl", "solvc"))
    if bonded:
        e.set_bonded(w.get("bonds"), w.get("bond_kr0"), w.get("angles"), w.get("angle_kt0"), w.get("dihedrals"), w.get("dihedral_prm"))
    if case.startswith("solvl"):
        e.set_thermostat(1, 300.0, 5.0, seed=7)
    if case.startswith("solvc"):
        e.set_thermostat(2, 300.0, 5.0, seed=7)
    e.set_option("halo_fused", 1 if halo == "fused" else 0)
    # the unbonded solvated system has very fast hydrogens; "adaptive" leaves the interval to the engine
    e.set_option("rebuild_every", 0 if sched == "adaptive" else int(os.environ.get("DD_EVERY", "5")) if case in ("lj", "ljx") else 2)
    e.set_option("dd_migrate", 0 if sched == "allgather" else 1)
    e.compute_forces()
    f0 = e.forces()
    en0 = e.energy()
    st0 = e.stats()
    press = virial = e_mols = 0.0
    if bonded:  # collective observables of a decomposed handle: pressure / virial and the between-molecules energy
        press, virial = e.pressure()
        e_mols = e.energy_between_mols(w["mol_id"])
    per_call = int(os.environ.get("DD_EXT", "0"))
    if per_call:
        e.set_option("defer_tail", int(os.environ.get("DD_DEFER", "1")))
        for k in range(n_steps // per_call):
            e.step(w["dt"], per_call, ext_forces_for_call(len(w["xyzq"]), k))
    else:
        e.step(w["dt"], n_steps)
    x = e.positions()
    v = e.velocities()
    ke_end = e.energy()["energy_kinetic"] if bonded else 0.0   # (collective: all ranks)
    st = e.stats()
    fused, why = e.halo_mode()
    interval, disp_frac = e.schedule()
    # rank-local asynchronous snapshot: owned atoms + their original ids
    own, gh = st["n_atoms"], st["n_ghosts"]
    sp, si = np.zeros((len(w["xyzq"]), 4), np.float32), np.full(len(w["xyzq"]), -1, np.int32)
    n_snap = e.snapshot_begin(sp, si)
    e.snapshot_wait()
    snap_ok = n_snap == own and np.array_equal(sp[:n_snap], x[si[:n_snap]])
    # the packed float3 hand-off, ids included and (same layout epoch) without
    s3, i3 = np.zeros((len(w["xyzq"]), 3), np.float32), np.full(len(w["xyzq"]), -1, np.int32)
    n3, ep = e.snapshot_begin_xyz(s3, i3)
    e.snapshot_wait()
    s3b = np.zeros_like(s3)
    n3b, ep2 = e.snapshot_begin_xyz(s3b)
    e.snapshot_wait()
    snap_ok = snap_ok and n3 == own and n3b == own and ep == ep2 and np.array_equal(i3[:n3], si[:n_snap]) and \
        np.array_equal(s3[:n3], x[i3[:n3], :3]) and np.array_equal(s3b[:n3], s3[:n3])
    if rank == 0:
        np.savez(out, ke_end=ke_end, e_bonded=en0["energy_potential_bonded"], pressure=press, virial=virial, e_mols=e_mols, fused=fused, why=why, snap_ok=snap_ok, interval=interval, disp_frac=disp_frac, f0=f0, x=x, v=v, e_pot=en0["energy_potential_nonbonded"], n_owned=st0["n_atoms"],
                 n_ghosts=st0["n_ghosts"], rebuilds=st["n_rebuilds"], violations=st["n_list_violations"],
                 ext_upload_bytes=st["ext_upload_bytes"])
    e.close()


if __name__ == "__main__":
    main()
