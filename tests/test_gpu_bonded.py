"""Bonded terms on the device against the fp64 oracle.  Needs a B200.

STATUS: bonded.cu was written after round 1's GPU budget was spent.  Its arithmetic is verified on the host
(tests/test_bonded_cpu.py compiles the same bonded_terms.h and checks it against the independent fp64 oracle and
finite differences); the kernel plumbing has not run on hardware yet, so these tests are allowed to fail without
turning the suite red (xfail, non-strict) until a GPU run has confirmed them."""
import numpy as np
import pytest

from molchanica_b200 import workloads as W
from util import FORCE_RTOL, trajectory_close

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="bonded.cu not yet run on hardware (round-1 GPU budget spent)")]


@pytest.fixture(scope="module")
def Engine():
    from molchanica_b200.engine import MdEngine
    return MdEngine


def test_bonded_forces_and_energies_match_oracle(Engine, oracle):
    w = W.bonded_globule(400)
    e = Engine.from_workload(w, bonded=True)
    e.compute_forces()
    f = e.forces()
    en = e.energy()
    nb = oracle.neighbors(w)
    f_nb, scale_nb, e_nb = oracle.forces(w, nb, precision=64)
    f_b, e_b = oracle.bonded(w)
    want = f_nb[:, :3] + f_b
    scale = scale_nb + np.abs(f_b).max(1)
    scale = np.maximum(scale, 1e-3 * scale.max())
    err = np.abs(f[:, :3].astype(np.float64) - want).max(1) / scale
    assert err.max() < 2 * FORCE_RTOL, err.max()
    assert np.allclose([en["energy_bond"], en["energy_angle"], en["energy_dihedral"]], e_b, rtol=2e-5)
    assert abs(en["energy_potential_bonded"] - e_b.sum()) < 2e-5 * e_b.sum()
    assert abs(en["energy_potential"] - (en["energy_potential_nonbonded"] + en["energy_potential_bonded"])) < 1e-9
    e.close()
    # without the bonded terms the same handle type gives the nonbonded forces only
    e = Engine.from_workload(w)
    e.compute_forces()
    assert e.energy()["energy_potential_bonded"] == 0.0
    e.close()


def test_flexible_water_box_follows_the_cpu_path(Engine, oracle):
    """C1 (216 flexible three-site waters, harmonic O-H and H-H bonds): 40 NVE steps on the GPU against the
    oracle run with the same bonds -- the configuration the reference can run on a CPU (BASELINE.json configs[0])."""
    w = W.water_box_c1()
    e = Engine.from_workload(w)
    e.set_bonded(w["bonds"], w["bond_kr0"])
    e.step(w["dt"], 40)
    ref = oracle.md_run(w, 40, precision=64, with_bonds=True)
    ok, worst, scale = trajectory_close(e.positions(), ref["xyzq"], w["xyzq"], w["box_ext"])
    assert ok, (worst, scale)
    en = e.energy()
    assert en["energy_bond"] > 0 and en["volume"] > 0 and 0.9 < en["density"] < 1.1
    e.close()


def test_bonded_terms_are_rejected_where_unsupported(Engine):
    w = W.bonded_globule(60, seed=77)
    e = Engine.from_workload(w)
    bad = np.array([[0, len(w["xyzq"])]], np.int32)
    with pytest.raises(Exception):
        e.set_bonded(bad, np.array([[100.0, 1.0]], np.float32))
    e.close()
