#!/usr/bin/env python
"""Measures the BASELINE.json configurations other than the bench headline (C2, C3, C5) on one GPU,
next to the CPU restatement, and prints one JSON object (kept under profiles/)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from molchanica_b200 import workloads as W  # noqa: E402
from molchanica_b200.engine import MdEngine  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

out = {}
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

# ---- C2: 1,231-atom globule in vacuum, 10k steps, batches of 10 like the GUI (src/md/mod.rs:45)
# (the synthetic globule has no bonded terms, so it is stepped with 0.2 fs to stay intact for 10k steps;
#  the cost of a step does not depend on dt -- ns/day is quoted at the reference default of 2 fs)
w = W.globule(temp_k=100.0)
e = MdEngine.from_workload(w)
DT_RUN = 0.0002
e.step(DT_RUN, 50)
t0 = time.perf_counter()
for _ in range(1000):
    e.step(DT_RUN, 10)
dt = time.perf_counter() - t0
st = e.stats()
out["C2"] = dict(atoms=len(w["xyzq"]), steps=10000, seconds=dt, steps_per_s=10000 / dt,
                 ns_per_day=10000 * w["dt"] * 1e-3 / dt * 86400, rebuilds=st["n_rebuilds"], launches=st["n_kernel_launches"])
t0 = time.perf_counter()
O.md_run(w, 200, precision=32)
c = time.perf_counter() - t0
out["C2"]["cpu_steps_per_s"] = 200 / c
out["C2"]["cpu_cores"] = O.num_threads()
e.close()

# ---- C3: 23,558 atoms solvated, pair-force and neighbour-build kernels individually
w = W.solvated_c3()
e = MdEngine.from_workload(w)
e.set_option("profiling", 1)
for _ in range(3):
    e.build_neighbors()
e.reset_timers()
for _ in range(10):
    e.build_neighbors()
sb = e.stats()
pair_ms = e.time_pair_kernel(reps=50, flush_l2=True)
n, p = len(w["xyzq"]), sb["n_pairs_listed"]
alg = 32.0 * n + 20.0 * p
out["C3"] = dict(atoms=n, pairs_listed=p, pair_kernel_ms=pair_ms, pair_alg_GBs=alg / pair_ms / 1e6, pair_frac_of_measured_hbm=alg / pair_ms / 1e6 / peak,
                 build_ms=sb["build_ms_sum"] / max(sb["builds_timed"], 1), build_alg_bytes=120.0 * n + 4.0 * p,
                 n_cells=sb["n_cells"])
nb = O.neighbors(w)
t0 = time.perf_counter()
for _ in range(3):
    O.forces(w, nb, precision=32)
out["C3"]["cpu_pair_ms"] = (time.perf_counter() - t0) / 3 * 1e3
t0 = time.perf_counter()
O.neighbors(w)
out["C3"]["cpu_build_ms"] = (time.perf_counter() - t0) * 1e3
out["C3"]["cpu_cores"] = O.num_threads()
e.close()

# ---- C5: docking scan, 10k poses x 5k receptor atoms x 40 ligand atoms
d = W.docking_c5()
e = MdEngine()
e.set_option("profiling", 1)
e.dock_score(d)
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    s = e.dock_score(d)
wall = (time.perf_counter() - t0) / reps
kms = e.last_dock_kernel_ms()
pairs = len(d["poses"]) * len(d["rec"]) * len(d["lig"])
out["C5"] = dict(poses=len(d["poses"]), receptor=len(d["rec"]), ligand=len(d["lig"]), pair_evals=pairs, kernel_ms=kms,
                 pair_evals_per_s_kernel=pairs / (kms * 1e-3), e2e_ms=wall * 1e3, poses_per_s_e2e=len(d["poses"]) / wall)
t0 = time.perf_counter()
O.dock_score(d, precision=32, poses=d["poses"][:500])
c = time.perf_counter() - t0
out["C5"]["cpu_poses_per_s"] = 500 / c
out["C5"]["cpu_cores"] = O.num_threads()
e.close()

# ---- the SURVEY 8f components (bonded terms, rigid 4-site water + virtual sites, Langevin, SPME); each in its own
# try block: they were written after round 1's GPU budget was spent and had not run on hardware when this was added
def _timed_steps(e, dt, n):
    e.step(dt, 20)
    t0 = time.perf_counter()
    e.step(dt, n)
    return n / (time.perf_counter() - t0)


try:  # C2 with its bonded terms at the reference's 2 fs
    w = W.bonded_globule(1231, seed=202)
    e = MdEngine.from_workload(w, bonded=True)
    out["C2_bonded"] = dict(atoms=len(w["xyzq"]), bonds=len(w["bonds"]), angles=len(w["angles"]), dihedrals=len(w["dihedrals"]),
                            steps_per_s=_timed_steps(e, 0.001, 2000), energy=e.energy())
    e.close()
except Exception as ex:  # noqa: BLE001
    out["C2_bonded"] = dict(error=str(ex))
try:  # OPC water box: SETTLE + virtual sites + SPME + Langevin, the reference's solvent set-up
    w = W.water_box_opc(m=12, L=37.3)
    e = MdEngine.from_workload(w)
    e.set_rigid_waters(w["rigid_waters"], w["d_oh"], w["d_hh"])
    e.set_virtual_sites(w["virtual_sites"], *w["vsite_ab"])
    e.set_pme(40, 40, 40)
    e.set_thermostat(1, 300.0, 1.0, seed=1)
    out["OPC_water"] = dict(atoms=len(w["xyzq"]), steps_per_s=_timed_steps(e, 0.002, 1000), energy=e.energy())
    e.close()
except Exception as ex:  # noqa: BLE001
    out["OPC_water"] = dict(error=str(ex))
try:  # the same box as NPT: CSVR + stochastic cell rescaling + drift removal (properties/crystal.rs:306-316)
    w = W.water_box_opc(m=12, L=37.3)
    e = MdEngine.from_workload(w)
    e.set_rigid_waters(w["rigid_waters"], w["d_oh"], w["d_hh"])
    e.set_virtual_sites(w["virtual_sites"], *w["vsite_ab"])
    e.set_pme(40, 40, 40)
    e.set_thermostat(2, 300.0, 10.0, seed=1)
    e.set_option("zero_com_drift", 100)
    e.set_barostat(2, 1.0, tau_ps=1.0, every=10, seed=3)
    rate = _timed_steps(e, 0.002, 1000)
    lo, hi = e.box()
    out["OPC_water_NPT"] = dict(atoms=len(w["xyzq"]), steps_per_s=rate, pressure_bar=e.pressure()[0], volume=float(np.prod(hi - lo)),
                                energy=e.energy())
    e.close()
except Exception as ex:  # noqa: BLE001
    out["OPC_water_NPT"] = dict(error=str(ex))
print(json.dumps(out))
