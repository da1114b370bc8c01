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
    ref = O.md_run(w, n_steps, precision=64, rigid_waters=(triples, D_OH, D_HH))
    ok, worst, sc = trajectory_close(x, ref["xyzq"], w["xyzq"], w["box_ext"])
    m = x[:, :3].astype(np.float64).reshape(-1, 3, 3)

    def d(a, b):
        v = a - b
        return np.linalg.norm(v - np.rint(v / ext) * ext, axis=1)
    geom = np.stack([d(m[:, 0], m[:, 1]), d(m[:, 0], m[:, 2]), d(m[:, 1], m[:, 2])], 1)
    res = dict(traj_ok=bool(ok), traj_worst=float(worst), geom_err=float(np.abs(geom - [D_OH, D_OH, D_HH]).max()),
               temperature=float(e.energy()["temperature"]))
    e.close()
    good = res["traj_ok"] and res["geom_err"] < 2e-5 and 100.0 < res["temperature"] < 600.0
    print(json.dumps(res))
    return 0 if good else 1


if __name__ == "__main__":
    sys.exit(main())
