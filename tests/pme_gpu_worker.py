"""On-device SPME check in a process of its own (spawned by tests/test_gpu_pme.py; pme.cu has not run on hardware
yet).  Prints one JSON line; exit code 0 = every check passed."""
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
from oracle import pme_oracle as P  # noqa: E402


def main():
    w = dict(W.water_box_c1(), coul_mode=2, alpha=0.35)
    K = (20, 20, 20)
    e = MdEngine.from_workload(w)
    e.compute_forces()
    f_off = e.forces()
    en_off = e.energy()
    e.set_pme(*K)
    e.compute_forces()
    f_on = e.forces()
    en_on = e.energy()
    e.close()
    # virial of the reciprocal sum + excluded-pair correction (mc_get_pressure): difference of the virials with and
    # without SPME (a flexible water, no constraints) against -dE/d(lambda) of the restatement under a uniform scaling
    e = MdEngine.from_workload(w)
    _, w_off = e.pressure()
    e.set_pme(*K)
    _, w_on = e.pressure()
    e.close()

    def u(lam):
        x = np.array(w["xyzq"], np.float64)
        x[:, :3] *= lam
        ext_l = np.asarray(w["box_ext"], np.float64) * lam
        lo_l = np.asarray(w["box_lo"], np.float64) * lam
        return P.spme(x, lo_l, ext_l, 0.35, K)[0] + P.excl_correction(x, ext_l, True, w["excl_start"], w["excl_idx"], 0.35)[0]
    h = 1e-3
    w_fd = -(u(1 + h) - u(1 - h)) / (2 * h)
    ext = np.asarray(w["box_ext"], np.float32)
    lo = np.asarray(w["box_lo"], np.float32)
    e_rec, f_rec = P.spme(w["xyzq"], lo, ext, 0.35, K)
    e_ex, f_ex = P.excl_correction(w["xyzq"], ext, True, w["excl_start"], w["excl_idx"], 0.35)
    e_pme = e_rec + e_ex + P.self_energy(w["xyzq"], 0.35)
    d = f_on[:, :3].astype(np.float64) - f_off[:, :3].astype(np.float64)
    scale = np.abs(f_rec + f_ex).max()
    res = dict(force_err=float(np.abs(d - (f_rec + f_ex)).max() / scale),
               energy_rel=float(abs(en_on["energy_pme"] - e_pme) / abs(e_pme)),
               nb_shift=float(abs((en_on["energy_potential_nonbonded"] - en_off["energy_potential_nonbonded"]) - en_on["energy_pme"])),
               off_is_zero=bool(en_off["energy_pme"] == 0.0), virial_pme=float(w_on - w_off), virial_fd=float(w_fd))
    # the fp32 forces are the sum of an fp32 real-space part (1e-5 of its own scale) and the reciprocal part
    good = res["force_err"] < 5e-4 and res["energy_rel"] < 5e-5 and res["nb_shift"] < 1e-6 * abs(e_pme) and res["off_is_zero"] and abs(res["virial_pme"] - w_fd) < 1e-4 * abs(w_fd)
    print(json.dumps(res))
    return 0 if good else 1


if __name__ == "__main__":
    sys.exit(main())
