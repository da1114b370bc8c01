"""Domain decomposition (slabs of cell layers along z + NCCL ghost exchange) against the oracle and
the single-GPU path.  Needs >= 2 B200s on the box (gpurun --gpus 2); skipped otherwise."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from molchanica_b200 import workloads as W
from util import FORCE_RTOL, energy_close, force_rel_err, trajectory_close

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except Exception:
        return 0


def _run(world, case, halo="fused", sched="fixed"):
    d = tempfile.mkdtemp()
    idf, out = os.path.join(d, "nccl_id"), os.path.join(d, "out.npz")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "dd_worker.py"), str(r), str(world), idf, case, out, halo, sched],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    logs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    return np.load(out)


@pytest.mark.parametrize("case,world,halo,sched", [
    ("lj", 2, "fused", "fixed"), ("lj", 2, "nccl", "fixed"), ("lj", 2, "fused", "allgather"), ("lj", 2, "fused", "adaptive"),
    ("solv", 2, "fused", "fixed"), ("solv", 2, "nccl", "fixed"),
    ("lj", 4, "fused", "fixed"), ("lj", 4, "fused", "adaptive"), ("lj", 8, "fused", "fixed"), ("lj", 8, "nccl", "fixed")])
def test_decomposed_run_matches_oracle(world, case, halo, sched, oracle):
    """solv (23.5k atoms, 14 A list radius) has only 4 cell layers along z: 2 ranks at most.
    halo = fused: ghosts are stored into the neighbours' arrays by kick_drift over mapped peer memory and
    the boundary rows' pair kernel waits on the flags (must really be active, not a silent NCCL fallback);
    halo = nccl: ncclSend / ncclRecv between the two kernels.
    sched = fixed: rebuild every 5 (2) steps, boundary layers migrate between neighbour ranks only;
    allgather: rebuilds all-gather the whole system; adaptive: the engine picks the interval from the
    largest displacement of the previous interval."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    sys.path.insert(0, HERE)
    from dd_worker import case_workload
    w, n_steps = case_workload(case, world)
    r = _run(world, case, halo, sched)
    assert int(r["fused"]) == (1 if halo == "fused" else 0), str(r["why"])
    nb = oracle.neighbors(w)
    f64, scale, en = oracle.forces(w, nb, precision=64)
    assert force_rel_err(r["f0"], f64, scale).max() < FORCE_RTOL
    assert energy_close(float(r["e_pot"]), en.sum(), f64[:, 3])
    ref = oracle.md_run(w, n_steps, precision=64)
    ok, worst, sc = trajectory_close(r["x"], ref["xyzq"], w["xyzq"], w["box_ext"])
    assert ok, (worst, sc)
    # (solv: free hydrogens of the unbonded box outrun skin/2 within its two-step schedule now and then; every such interval
    # is counted since the rebuild table carries the largest displacement -- the trajectory above is still the oracle's)
    assert int(r["violations"]) == 0 or (case == "solv" and int(r["violations"]) <= int(r["rebuilds"]))
    if sched == "adaptive":
        assert int(r["interval"]) >= 10 and int(r["rebuilds"]) >= 2 and 0.0 < float(r["disp_frac"]) < 1.0
    assert bool(r["snap_ok"]), "rank-local snapshot differs from the gathered positions"
    assert int(r["n_owned"]) < len(w["xyzq"]) and int(r["n_ghosts"]) > 0


def _single_handle_reference(w, n_steps, thermostat):
    """The same system on one handle of the same library: initial forces, energies, positions after n_steps.
    thermostat: 0 none, 1 Langevin, 2 CSVR (bools of older callers: True = Langevin)."""
    from molchanica_b200.engine import MdEngine
    e = MdEngine.from_workload(w, bonded=True)
    if thermostat:
        e.set_thermostat(int(thermostat), 300.0, 5.0, seed=7)
    e.set_option("rebuild_every", 2)
    e.compute_forces()
    f0, en = e.forces(), e.energy()
    en["pressure"], en["virial"] = e.pressure()
    en["between_mols"] = e.energy_between_mols(w["mol_id"])
    e.step(w["dt"], n_steps)
    x = e.positions()
    en["ke_end"] = e.energy()["energy_kinetic"]
    e.close()
    return f0, en, x


def check_bonded_decomposed(r, w, n_steps, thermostat):
    """Shared with tests/test_library_on_host.py: a decomposed run with bonded terms against the single-handle run."""
    f0, en, x = _single_handle_reference(w, n_steps, thermostat)
    # forces: same terms, fp32 atomics in another order -> a few ulp of the largest contribution per atom
    scale = np.abs(f0[:, :3]).max(1) + 1e-2 * np.abs(f0[:, :3]).max()
    assert (np.abs(r["f0"][:, :3] - f0[:, :3]).max(1) / scale).max() < 2e-5
    assert abs(float(r["e_bonded"]) - en["energy_potential_bonded"]) < 1e-5 * abs(en["energy_potential_bonded"]) + 1e-3
    assert abs(float(r["e_pot"]) - en["energy_potential_nonbonded"]) < 1e-5 * abs(en["energy_potential_nonbonded"]) + 1e-3
    assert abs(float(r["virial"]) - en["virial"]) < 2e-5 * abs(en["virial"]) + 1e-2, (float(r["virial"]), en["virial"])
    assert abs(float(r["pressure"]) - en["pressure"]) < 2e-5 * abs(en["pressure"]) + 1.0, (float(r["pressure"]), en["pressure"])
    assert abs(float(r["e_mols"]) - en["between_mols"]) < 1e-5 * abs(en["between_mols"]) + 1e-3, (float(r["e_mols"]), en["between_mols"])
    d = r["x"][:, :3] - x[:, :3]
    d -= np.rint(d / w["box_ext"]) * w["box_ext"]
    assert np.abs(d).max() < 2e-4, np.abs(d).max()
    # the kinetic energy after the steps: a thermostat fed with one rank's share of it would scale differently
    assert abs(float(r["ke_end"]) - en["ke_end"]) < 2e-5 * en["ke_end"], (float(r["ke_end"]), en["ke_end"])
    assert bool(r["snap_ok"]) and int(r["n_owned"]) < len(w["xyzq"]) and int(r["n_ghosts"]) > 0 and int(r["rebuilds"]) >= 2


@pytest.mark.parametrize("case,halo", [("solvb", "fused"), ("solvb", "nccl"), ("solvl", "fused"), ("solvc", "fused")])
def test_bonded_terms_and_langevin_on_a_decomposed_handle(case, halo):
    """SURVEY 8f row 3 across slab boundaries: bonds, angles and dihedrals (with their exclusions and 1-4 pairs) on two ranks
    -- every rank evaluates the terms that touch its owned atoms, partners are ghosts, energies are shared out by owned atoms
    and all-reduced -- and the Langevin thermostat, whose noise is keyed by (seed, step, original id) and therefore the same on
    any decomposition; solvc: the CSVR thermostat, whose one scaling factor per step is a function of (seed, step, kinetic
    energy) -- the kinetic energy is all-reduced on the stream inside the step.  Reference: the single-handle run of the same library (itself checked against the oracle in
    tests/test_gpu_bonded.py / test_gpu_langevin.py)."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, HERE)
    from dd_worker import case_workload
    w, n_steps = case_workload(case, 2)
    r = _run(2, case, halo, "fixed")
    assert int(r["fused"]) == (1 if halo == "fused" else 0), str(r["why"])
    check_bonded_decomposed(r, w, n_steps, 1 if case.startswith("solvl") else (2 if case.startswith("solvc") else 0))
