"""Known-answer tests of the oracle that do not depend on any engine: analytic
two-body values, Newton's third law, list == brute force, NVE energy conservation on C1."""
import numpy as np

from molchanica_b200 import workloads as W


def _two_atoms(r, sigma=3.4, eps=0.24, q=(0.0, 0.0), coul_mode=0):
    xyzq = np.array([[0, 0, 0, q[0]], [r, 0, 0, q[1]]], np.float32)
    return dict(xyzq=xyzq, vel=np.zeros((2, 4), np.float32), type=np.zeros(2, np.uint16),
                ljtab=np.array([[[sigma, eps]]], np.float32), box_lo=np.zeros(3, np.float32),
                box_ext=np.full(3, 100.0, np.float32), periodic=False, rc_lj=50.0, rc_q=50.0, skin=1.0,
                coul_mode=coul_mode, excl_start=None, excl_idx=None, pairs14=None, scale14_lj=0.5, scale14_q=1 / 1.2)


def test_lj_minimum_and_zero_crossing(oracle):
    sigma, eps = 3.4, 0.24
    w = _two_atoms(2 ** (1 / 6) * sigma, sigma, eps)
    f, _, en = oracle.forces(w, oracle.neighbors(w), precision=64)
    assert abs(en[0] + eps) < 1e-6 and np.abs(f[:, :3]).max() < 1e-5      # F = 0, E = -eps at r_min
    w = _two_atoms(sigma, sigma, eps)
    f, _, en = oracle.forces(w, oracle.neighbors(w), precision=64)
    assert abs(en[0]) < 1e-6                                               # E = 0 at r = sigma
    assert abs(abs(f[0, 0]) - 24 * eps / sigma) < 1e-5 * 24 * eps / sigma  # |F| = 24 eps / sigma
    assert f[0, 0] < 0 < f[1, 0]                                           # repulsive: pushes the atoms apart


def test_coulomb_pair_in_kcal_per_mol(oracle):
    s = W.COULOMB_SCALE
    w = _two_atoms(5.0, 1.0, 0.0, q=(0.5 * s, -0.4 * s), coul_mode=1)
    f, _, en = oracle.forces(w, oracle.neighbors(w), precision=64)
    assert abs(en[1] - 332.0522 * 0.5 * -0.4 / 5.0) < 1e-4
    assert abs(f[1, 0] - (-332.0522 * 0.2 / 25.0)) < 1e-4 and f[0, 0] > 0  # opposite charges attract


def test_newton_third_law_and_list_equivalence(oracle):
    for w in (W.lj_fluid(m=7), W.water_box_c1(), W.globule(400, seed=9)):
        nb = oracle.neighbors(w)
        nb_b = oracle.neighbors(w, brute=True)
        assert np.array_equal(nb[0], nb_b[0]) and np.array_equal(nb[1], nb_b[1])
        f, sa, _ = oracle.forces(w, nb, precision=64)
        assert np.abs(f[:, :3].astype(np.float64).sum(0)).max() < 1e-4 * float(sa.max())
        # symmetry of the full list
        rows = {(i, int(j)) for i in range(len(nb[0]) - 1) for j in nb[1][nb[0][i]:nb[0][i + 1]]}
        assert all((j, i) in rows for (i, j) in list(rows)[:5000])


def test_c1_water_box_nve_plumbing(oracle):
    """BASELINE config 1: 216-water box, NVE on the CPU (erfc real-space Coulomb so the energy is
    continuous at the cutoff; flexible harmonic bonds in the oracle only)."""
    w = W.water_box_c1()
    r = oracle.md_run(w, 300, precision=32, want_energies=True, with_bonds=True)
    tot = r["energies"].sum(1)
    ke0 = r["energies"][0, 3]
    assert abs(tot[-1] - tot[0]) < 0.02 * ke0, (tot[0], tot[-1])
    assert r["rebuilds"] >= 1
