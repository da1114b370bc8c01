"""On-device Langevin thermostat check in a process of its own (spawned by tests/test_gpu_langevin.py; the O-step
kernel has not run on hardware yet).  Prints one JSON line; exit code 0 = every check passed."""
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


def main():
    # same noise on both sides (Philox keyed by seed, atom id, step): the trajectories must agree
    w = W.lj_fluid(m=12)
    e = MdEngine.from_workload(w)
    e.set_thermostat(1, 150.0, 10.0, seed=42)
    e.step(w["dt"], 25)
    ref = O.md_run(w, 25, precision=64, langevin=(150.0, 10.0, 42))
    ok, worst, sc = trajectory_close(e.positions(), ref["xyzq"], w["xyzq"], w["box_ext"])
    dv = float(np.abs(e.velocities()[:, :3] - ref["vel"][:, :3]).max())
    e.close()
    # thermalisation: 40 K -> 120 K
    w = W.lj_fluid(m=10, temp_k=40.0)
    e = MdEngine.from_workload(w)
    e.set_thermostat(1, 120.0, 20.0, seed=5)
    e.step(w["dt"], 300)
    temps = []
    for _ in range(10):
        e.step(w["dt"], 20)
        e.compute_forces()
        temps.append(e.energy()["temperature"])
    e.close()
    # CSVR: the scale factor is a function of (seed, step, kinetic energy) -> same trajectory as the oracle
    w = W.lj_fluid(m=12)
    e = MdEngine.from_workload(w)
    e.set_thermostat(2, 150.0, 20.0, seed=9)
    e.step(w["dt"], 25)
    ref = O.md_run(w, 25, precision=64, csvr=(150.0, 20.0, 9))
    ok2, worst2, _ = trajectory_close(e.positions(), ref["xyzq"], w["xyzq"], w["box_ext"])
    dv2 = float(np.abs(e.velocities()[:, :3] - ref["vel"][:, :3]).max())
    e.close()
    res = dict(traj_ok=bool(ok), traj_worst=float(worst), dv=dv, temperature=float(np.mean(temps)), csvr_traj_ok=bool(ok2),
               csvr_traj_worst=float(worst2), csvr_dv=dv2)
    good = (res["traj_ok"] and res["dv"] < 5e-3 and abs(res["temperature"] - 120.0) < 12.0 and res["csvr_traj_ok"] and
            res["csvr_dv"] < 5e-3)
    print(json.dumps(res))
    return 0 if good else 1


if __name__ == "__main__":
    sys.exit(main())
