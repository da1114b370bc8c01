"""Docking pose-energy scan vs the oracle (fp64 truth) and the golden fixture."""
import os

import numpy as np
import pytest

from molchanica_b200 import workloads as W

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "md_small.npz")


def _check(gpu, ref):
    # columns: score, vdw, hydrophobic, electrostatic, coulomb_e.  vdw spans ~15 orders of magnitude
    # over clashing poses, so every column is compared relative to its own magnitude per pose, with
    # an absolute floor for poses whose sum cancels.
    for col, floor in ((1, 1e-3), (2, 1e-4), (3, 1e-3), (4, 2e-2)):
        tol = 2e-5 * np.abs(ref[:, col]) + floor
        assert np.all(np.abs(gpu[:, col] - ref[:, col]) <= tol), (col, np.abs(gpu[:, col] - ref[:, col]).max())
    score_ref = ref[:, 1] + ref[:, 2] + 10.0 * ref[:, 3]
    assert np.all(np.abs(gpu[:, 0] - score_ref) <= 2e-5 * np.abs(score_ref) + 2e-2)


def test_dock_scan_matches_oracle():
    from molchanica_b200.engine import MdEngine
    from oracle import oracle_py as O
    d = W.docking_c5(n_rec=1500, n_lig=24, n_poses=300, seeds=(525, 526, 527))
    e = MdEngine()
    _check(e.dock_score(d), O.dock_score(d, precision=64))
    e.close()


def test_dock_scan_matches_golden():
    from molchanica_b200.engine import MdEngine
    g = np.load(GOLD)
    d = {k.split(".", 1)[1]: g[k] for k in g.files if k.startswith("dock.")}
    e = MdEngine()
    _check(e.dock_score(d), d["scores64"])
    e.close()


def test_dock_scan_full_size_is_invariant_under_pose_order():
    """C5 at full size (10k poses x 5k receptor atoms): a permutation of the poses permutes the
    scores (each pose is an independent unit of work)."""
    from molchanica_b200.engine import MdEngine
    d = W.docking_c5()
    e = MdEngine()
    s = e.dock_score(d)
    perm = np.random.default_rng(3).permutation(len(d["poses"]))
    s2 = e.dock_score(d, poses=d["poses"][perm])
    assert np.array_equal(s[perm], s2)
    assert np.isfinite(s[:, 1:]).all()
    e.close()
