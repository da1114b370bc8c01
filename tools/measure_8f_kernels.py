#!/usr/bin/env python
"""Workload for the per-kernel timing of the SURVEY 8f components (run under ncu, see tools/gpu_r2_u.sh):
 (1) a 55,296-atom OPC water box (24^3 four-site waters, 74.6 A): SETTLE + RATTLE + virtual sites + SPME (72^3) + Langevin, NVT;
 (2) the same box with CSVR + stochastic cell rescaling (NPT);
 (3) a 6,000-atom bonded globule (bonds, angles, dihedrals, exclusions, 1-4 pairs) in vacuum."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from molchanica_b200 import workloads as W  # noqa: E402
from molchanica_b200.engine import MdEngine  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
M = int(os.environ.get("MC_8F_M", "24"))  # waters per box edge (host-build dry runs: 6)
w = W.water_box_opc(m=M, L=74.6 * M / 24)
for mode in ("nvt", "npt"):
    e = MdEngine.from_workload(w)
    e.set_rigid_waters(w["rigid_waters"], w["d_oh"], w["d_hh"])
    e.set_virtual_sites(w["virtual_sites"], *w["vsite_ab"])
    e.set_pme(3 * M, 3 * M, 3 * M)
    if mode == "nvt":
        e.set_thermostat(1, 300.0, 1.0, seed=1)
    else:
        e.set_thermostat(2, 300.0, 10.0, seed=1)
        e.set_option("zero_com_drift", 10)
        e.set_barostat(2, 1.0, tau_ps=1.0, every=10, seed=3)
    e.step(0.002, steps)
    en = e.energy()
    print(mode, len(w["xyzq"]), "atoms", {k: round(float(v), 3) for k, v in en.items() if isinstance(v, (int, float))}, flush=True)
    e.close()
w = W.bonded_globule(int(os.environ.get("MC_8F_GLOBULE", "6000")), seed=202)
e = MdEngine.from_workload(w, bonded=True)
e.step(0.001, steps)
print("bonded", len(w["xyzq"]), "atoms", len(w["bonds"]), "bonds", len(w["angles"]), "angles", len(w["dihedrals"]), "dihedrals", flush=True)
e.close()
