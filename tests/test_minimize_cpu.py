"""The minimiser behind mc_minimize_energy (md.minimize_energy of the reference, ui/mol_editor.rs:375), restated in fp64
(oracle.minimize): energies never go up, a jittered LJ crystal relaxes towards its lattice energy, static atoms stay."""
import numpy as np

from molchanica_b200 import workloads as W


def test_oracle_minimiser_descends_and_respects_static_atoms(oracle):
    w = W.lj_fluid(m=6, temp_k=0.0)                 # 216 argon atoms on the jittered lattice
    flags = np.zeros(len(w["xyzq"]), np.uint8)
    flags[::10] = 1
    w = dict(w, flags=flags, pairs14=None)
    r = oracle.minimize(w, 60)
    e = r["energies"]
    assert r["accepted"] >= 10 and np.all(np.diff(e) <= 0) and r["e_final"] < r["e_initial"] - 1.0
    assert np.array_equal(r["xyzq"][::10], w["xyzq"][::10])
    moved = np.abs(r["xyzq"][:, :3] - w["xyzq"][:, :3]).max(1)
    assert moved[flags == 0].max() > 1e-3
