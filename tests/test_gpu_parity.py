"""Parity of the CUDA path (through the C ABI) with the oracle: neighbour indices bit-exact,
forces within 1e-5 of the fp64 truth relative to sum|f_ij|, energies within 1e-5, short
trajectories, and the committed golden fixtures.  Needs a B200."""
import os

import numpy as np
import pytest

from molchanica_b200 import workloads as W
from util import FORCE_RTOL, energy_close, force_rel_err, numpy_row, trajectory_close

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "md_small.npz")


@pytest.fixture(scope="module")
def Engine():
    from molchanica_b200.engine import MdEngine
    return MdEngine


def _cases():
    return {
        "lj1728": lambda: W.lj_fluid(m=12),
        "lj8000": lambda: W.lj_fluid(m=20),
        "water648": W.water_box_c1,
        "glob1231": W.globule,
        "solv23558": W.solvated_c3,
    }


@pytest.mark.parametrize("name", list(_cases()))
def test_neighbour_list_bit_exact_and_forces(name, Engine, oracle):
    w = _cases()[name]()
    e = Engine.from_workload(w)
    e.build_neighbors()
    start, idx = e.neighbors()
    o_start, o_idx = oracle.neighbors(w)
    assert np.array_equal(start, o_start), "row lengths differ"
    assert np.array_equal(idx, o_idx), "neighbour indices differ"
    assert e.stats()["n_pairs_listed"] == len(o_idx)

    e.compute_forces()
    f = e.forces()
    f64, sumabs, en = oracle.forces(w, (o_start, o_idx), precision=64)
    err = force_rel_err(f, f64, sumabs)
    assert err.max() < FORCE_RTOL, f"max force error {err.max():.3e} at atom {err.argmax()}"
    assert energy_close(e.energy()["energy_potential_nonbonded"], en.sum(), f64[:, 3])
    # per-atom energy rows
    assert np.abs(f[:, 3] - f64[:, 3]).max() < 1e-5 * max(1.0, float(np.abs(f64[:, 3]).max()))
    e.close()


@pytest.mark.parametrize("lanes", [4, 8, 16, 32])
def test_every_lane_width_gives_the_same_forces(lanes, Engine, oracle):
    w = W.solvated_c3()
    e = Engine.from_workload(w)
    e.set_option("pair_lanes", lanes)
    e.compute_forces()
    f = e.forces()
    nb = oracle.neighbors(w)
    f64, sumabs, _ = oracle.forces(w, nb, precision=64)
    assert force_rel_err(f, f64, sumabs).max() < FORCE_RTOL
    e.close()


def test_overrides_isolate_lj_and_coulomb(Engine, oracle):
    """MdOverrides.lj_disabled / coulomb_disabled (reference src/md/mod.rs:671-686)."""
    w = W.globule()
    nb = oracle.neighbors(w)
    e = Engine.from_workload(w)
    for lj_off, q_off in ((True, False), (False, True)):
        e.set_overrides(lj_off, q_off)
        e.compute_forces()
        f64, sumabs, en = oracle.forces(w, nb, precision=64, lj_on=not lj_off, coul_on=not q_off)
        assert force_rel_err(e.forces(), f64, sumabs).max() < FORCE_RTOL
        assert energy_close(e.energy()["energy_potential_nonbonded"], en.sum(), f64[:, 3])
    e.close()


def test_erfc_real_space_mode(Engine, oracle):
    w = W.solvated_c3()
    w["coul_mode"] = 2
    e = Engine.from_workload(w)
    e.compute_forces()
    nb = oracle.neighbors(w)
    f64, sumabs, en = oracle.forces(w, nb, precision=64)
    assert force_rel_err(e.forces(), f64, sumabs).max() < FORCE_RTOL
    assert energy_close(e.energy()["energy_potential_nonbonded"], en.sum(), f64[:, 3])
    e.close()


@pytest.mark.parametrize("name", ["lj1728", "glob1231", "water648"])
def test_short_trajectory_follows_the_cpu_path(name, Engine, oracle):
    w = _cases()[name]()
    n_steps = 20 if name.startswith("lj") else 6
    e = Engine.from_workload(w)
    e.step(w["dt"], n_steps)
    ref = oracle.md_run(w, n_steps, precision=64)
    ok, worst, scale = trajectory_close(e.positions(), ref["xyzq"], w["xyzq"], w["box_ext"] if w["periodic"] else None)
    assert ok, (worst, scale)
    v = e.velocities()
    verr = np.abs(v[:, :3] - ref["vel"][:, :3]).max(1)
    assert np.quantile(verr, 0.99) < 2e-4 * max(float(np.abs(ref["vel"][:, :3]).max()), 1.0)
    assert e.stats()["n_steps"] == n_steps
    e.close()


def test_async_snapshots_match_blocking_reads(Engine):
    """mc_snapshot_begin / mc_snapshot_wait (the Snapshot queue of reference src/md/mod.rs:118-152): two
    snapshots in flight while the integrator keeps stepping; each equals the blocking read of its step."""
    w = W.lj_fluid(m=16)
    e = Engine.from_workload(w)
    n = len(w["xyzq"])
    bufs = [np.zeros((n, 4), np.float32) for _ in range(3)]
    want = []
    e.step(w["dt"], 3)
    want.append(e.positions())
    assert e.snapshot_begin(bufs[0]) == n
    e.step(w["dt"], 2)
    want.append(e.positions())
    e.snapshot_begin(bufs[1])
    e.step(w["dt"], 4)
    e.snapshot_wait()
    assert np.array_equal(bufs[0], want[0])
    want.append(e.positions())
    e.snapshot_begin(bufs[2])
    e.snapshot_wait()
    e.snapshot_wait()
    assert np.array_equal(bufs[1], want[1]) and np.array_equal(bufs[2], want[2])
    # positions + velocities (Snapshot.atom_velocities), also right after a pipelined call with external forces
    ext = np.zeros((n, 3), np.float32)
    ext[::5, 1] = 2.0
    e.step(w["dt"], 2, ext_forces=ext)
    px, pv = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
    assert e.snapshot_begin_pv(px, pv) == n
    e.snapshot_wait()
    assert np.array_equal(px, e.positions()) and np.array_equal(pv, e.velocities())
    e.close()


@pytest.mark.parametrize("opt", [("subcell_sort", 1), ("subcell_sort", 0), ("sync_rebuild", 1), ("rebuild_every", 4)])
def test_build_and_rebuild_policies_keep_parity(opt, Engine, oracle):
    """The Morton sub-cell sort key, the host-published displacement flag (default), the synchronous flag
    read and the fixed rebuild schedule all list the oracle's pairs and follow the oracle's trajectory."""
    w = W.lj_fluid(m=14, temp_k=400.0)  # hot enough that the displacement criterion fires within 60 steps
    e = Engine.from_workload(w)
    e.set_option(*opt)
    e.build_neighbors()
    start, idx = e.neighbors()
    o_start, o_idx = oracle.neighbors(w)
    assert np.array_equal(start, o_start) and np.array_equal(idx, o_idx)
    e.step(w["dt"], 60)
    assert e.stats()["n_rebuilds"] >= 3
    ref = oracle.md_run(w, 60, precision=64)
    ok, worst, scale = trajectory_close(e.positions(), ref["xyzq"], w["xyzq"], w["box_ext"])
    assert ok, (worst, scale)
    e.close()


def test_external_forces_and_static_atoms(Engine, oracle):
    """step(dev, dt, Some(forces)) (reference src/mol_alignment.rs:346) and AtomDynamics.static_."""
    w = W.globule(300, seed=33)
    rng = np.random.default_rng(5)
    ext = rng.normal(0, 5.0, (len(w["xyzq"]), 3)).astype(np.float32)
    e = Engine.from_workload(w)
    e.step(w["dt"], 5, ext_forces=ext)
    ref = oracle.md_run(w, 5, precision=64, ext_force=ext)
    ok, worst, scale = trajectory_close(e.positions(), ref["xyzq"], w["xyzq"])
    assert ok, (worst, scale)
    e.close()
    flags = np.zeros(len(w["xyzq"]), np.uint8)
    flags[::3] = 1
    w2 = dict(w, flags=flags)
    e = Engine.from_workload(w2)
    e.step(w["dt"], 5)
    x = e.positions()
    assert np.array_equal(x[::3], w["xyzq"][::3])
    assert np.abs(x[1::3, :3] - w["xyzq"][1::3, :3]).max() > 0
    e.close()


def _run_with_changing_ext(Engine, w, defer, n_calls, steps_per_call, poke):
    """One mc_step(dt, k, ext) per call with a different array every call, positions read back after every call (as the
    reference's alignment loop does, src/mol_alignment.rs:318-353).  `poke` = observers / setters thrown in on the way."""
    n = len(w["xyzq"])
    e = Engine.from_workload(w)
    e.set_option("defer_tail", 1 if defer else 0)
    rng = np.random.default_rng(77)
    seen = []
    for k in range(n_calls):
        ext = np.zeros((n, 3), np.float32)
        ext[k % 7::7] = rng.normal(0, 3.0, ext[k % 7::7].shape)
        e.step(w["dt"], steps_per_call, ext_forces=ext)
        ext[:] = np.nan                               # the array may be reused the moment the call returns
        seen.append(e.positions())
        if poke and k == 3:
            seen.append(e.velocities())               # observer in the middle: closes the open half kick
        if poke and k == 6:
            seen.append(e.energy()["energy_kinetic"])
        if poke and k == 9:
            v = e.velocities()
            v[:, :3] *= 0.5
            e.set_velocities(v)                       # setter in the middle: must see, then replace, the finished velocities
        if poke and k == 12:
            e.step(w["dt"], 2)                        # a call without external forces in between
    out = dict(x=e.positions(), v=e.velocities(), f=e.forces(), en=e.energy(), seen=seen, rebuilds=e.stats()["n_rebuilds"],
               steps=e.stats()["n_steps"])
    e.close()
    return out


@pytest.mark.parametrize("steps_per_call,poke,skin", [(1, False, 0.35), (1, True, 0.35), (3, True, 0.35), (1, True, 1.5), (3, False, 1.5)])
def test_pipelined_external_forces_are_invisible_through_the_abi(steps_per_call, poke, skin, Engine, oracle):
    """mc_step with external forces returns after the last drift and finishes that step (force evaluation + second half
    kick) under the upload of the next call's array (engine.cu, `defer_tail`).  Through the ABI that must be invisible:
    bit-identical positions, velocities, forces and energies with the option off, whatever is called in between, as long
    as no rebuild falls into the run (a rebuild changes the summation order of a row; with the option on, both half kicks
    around it use the forces of the NEW list, with it off the first uses the old one: a last-bit difference); and the
    trajectory is the oracle's."""
    w = dict(W.lj_fluid(m=12), skin=skin)             # small skin: rebuilds fall inside the run
    maxwell = np.random.default_rng(3).normal(0, 1.0, (len(w["xyzq"]), 3)).astype(np.float32)
    w["vel"] = w["vel"].copy()
    w["vel"][:, :3] += 2.0 * maxwell * w["vel"][:, 3:4] ** 0.5
    a = _run_with_changing_ext(Engine, w, True, 16, steps_per_call, poke)
    b = _run_with_changing_ext(Engine, w, False, 16, steps_per_call, poke)
    assert a["steps"] == b["steps"]
    if skin < 1.0:
        assert a["rebuilds"] >= 2 or poke            # (the halved velocities of the poked run may avoid the second one)
        same = lambda u, v: np.allclose(u, v, rtol=2e-5, atol=2e-5)
    else:
        assert a["rebuilds"] == 1
        same = np.array_equal
    for u, v in zip(a["seen"], b["seen"]):
        assert same(np.asarray(u), np.asarray(v))
    assert same(a["x"], b["x"]) and same(a["v"], b["v"]) and same(a["f"], b["f"])
    assert all(same(np.float64(a["en"][k]), np.float64(b["en"][k])) for k in a["en"])
    if not poke:
        # against the oracle: the same arrays, call by call
        n = len(w["xyzq"])
        rng = np.random.default_rng(77)
        cur = dict(w)
        for k in range(16):
            ext = np.zeros((n, 3), np.float32)
            ext[k % 7::7] = rng.normal(0, 3.0, ext[k % 7::7].shape)
            r = oracle.md_run(cur, steps_per_call, precision=64, ext_force=ext)
            cur = dict(cur, xyzq=r["xyzq"], vel=r["vel"])
        ok, worst, scale = trajectory_close(a["x"], cur["xyzq"], w["xyzq"], w["box_ext"])
        assert ok, (worst, scale)


def _pair_virial64(w, nbr, coul_mode):
    """fp64 virial sum_{i<j} r_ij . f_ij of the listed pairs inside the cutoffs, written out in numpy independently of the
    device code: LJ 24 eps (2 s^12 - s^6); Coulomb qq/r (plain) or qq (erfc(ar)/r + 2a/sqrt(pi) exp(-a^2 r^2)) (Ewald real space)."""
    from scipy.special import erfc
    start, idx = nbr
    x = np.asarray(w["xyzq"], np.float64)
    ext = np.asarray(w["box_ext"], np.float64)
    i = np.repeat(np.arange(len(x)), np.diff(start))
    d = x[i, :3] - x[idx, :3]
    if w["periodic"]:
        d -= ext * np.rint(d / ext)
    r2 = (d * d).sum(1)
    tab = np.asarray(w["ljtab"], np.float64)
    t = np.asarray(w["type"])
    sig, eps = tab[t[i], t[idx], 0], tab[t[i], t[idx], 1]
    s6 = (sig * sig / r2) ** 3
    wl = np.where(r2 < float(np.float32(w["rc_lj"])) ** 2, 24.0 * eps * s6 * (2.0 * s6 - 1.0), 0.0)
    qq = x[i, 3] * x[idx, 3]
    r = np.sqrt(r2)
    a = float(w.get("alpha", 0.35))
    wq = qq / r if coul_mode == 1 else qq * (erfc(a * r) / r + 2.0 * a / np.sqrt(np.pi) * np.exp(-a * a * r2))
    wq = np.where((r2 < float(np.float32(w["rc_q"])) ** 2) & (coul_mode != 0), wq, 0.0)
    return 0.5 * float(wl.sum() + wq.sum())


def test_pressure_of_an_lj_fluid(Engine, oracle):
    """mc_get_pressure (SnapshotEnergyData.pressure): virial of the listed pairs against the fp64 sum, P = (2 KE + W) / 3V."""
    w = W.lj_fluid(m=12)
    e = Engine.from_workload(w)
    e.step(w["dt"], 20)                                 # off the lattice
    x, v = e.positions(), e.velocities()
    p_bar, vir = e.pressure()
    e.close()
    ws = dict(w, xyzq=x)
    w64 = _pair_virial64(ws, oracle.neighbors(ws), 0)
    assert abs(vir - w64) < 2e-5 * abs(w64), (vir, w64)
    ke = 0.5 * float(((v[:, :3].astype(np.float64) ** 2).sum(1) / v[:, 3]).sum()) / 418.4
    vol = float(np.prod(np.asarray(w["box_ext"], np.float64)))
    assert abs(p_bar - (2 * ke + w64) / (3 * vol) * 69476.95) < 2e-5 * (abs(p_bar) + 2 * ke / (3 * vol) * 69476.95)
    # sanity of the magnitude: a dense LJ liquid a few steps off its lattice sits within a few kbar of zero
    assert abs(p_bar) < 6000.0


@pytest.mark.parametrize("coul_mode", [1, 2])
def test_pressure_with_charges_bonds_and_exclusions(coul_mode, Engine, oracle):
    """Flexible water (harmonic O-H / H-H bonds, intramolecular pairs excluded): pair virial (LJ + plain or Ewald real-space
    Coulomb) + bonded virial; the bonded part against -dU/d(lambda) of the oracle's bonded energy under a uniform scaling."""
    w = dict(W.water_box_c1(), coul_mode=coul_mode, alpha=0.35)
    e = Engine.from_workload(w)
    e.set_bonded(w["bonds"], w["bond_kr0"])
    e.step(0.0005, 10)
    x = e.positions()
    _, vir = e.pressure()
    e.close()
    ws = dict(w, xyzq=x)
    w_pair = _pair_virial64(ws, oracle.neighbors(ws), coul_mode)

    def u(lam):
        xs = np.array(x, np.float64)
        xs[:, :3] *= lam
        return float(np.sum(oracle.bonded(dict(ws, xyzq=xs.astype(np.float32), box_ext=np.asarray(w["box_ext"], np.float64) * lam))[1]))
    h = 1e-3
    w_bond = -(u(1 + h) - u(1 - h)) / (2 * h)
    assert abs(w_bond) > 1.0
    scale = abs(w_pair) + abs(w_bond)
    assert abs(vir - (w_pair + w_bond)) < 5e-4 * scale, (vir, w_pair, w_bond)


def _free_rotors(w, seed):
    """Random thermal velocities for the atoms of a workload whose interactions are switched off."""
    rng = np.random.default_rng(seed)
    v = w["vel"].copy()
    v[:, :3] = rng.normal(0, 1.0, (len(v), 3)) * np.sqrt(0.0019872041 * 300.0 * 418.4 * v[:, 3:4])
    return v


@pytest.mark.parametrize("kind", ["settle", "shake"])
def test_constraint_virial_of_free_rigid_rotors(kind, Engine):
    """Known answer for the constraint part of mc_get_pressure: molecules that do not interact at all.  The only forces are
    the constraint forces that keep a rotating rigid body together (centripetal: W_c = -2 KE_rot), so 2 KE + W must be
    twice the kinetic energy of the centres of mass -- the ideal-gas pressure of N molecules, not of 3 N atoms."""
    w = W.water_box_c1()
    n = len(w["xyzq"])
    if kind == "shake":                                   # rigid O-H diatomics: drop the second hydrogen
        keep = np.arange(n).reshape(-1, 3)[:, :2].ravel()
        w = dict(w, xyzq=w["xyzq"][keep], vel=w["vel"][keep], type=w["type"][keep], excl_start=None, excl_idx=None)
        n, per = len(keep), 2
    else:
        w = dict(w, excl_start=None, excl_idx=None)
        per = 3
    w["vel"] = _free_rotors(w, 12)
    dt = 0.00025
    e = Engine.from_workload(w)
    e.set_overrides(lj_disabled=True, coulomb_disabled=True)
    ids = np.arange(n, dtype=np.int32).reshape(-1, per)
    if kind == "settle":
        e.set_rigid_waters(ids, 0.9572, 1.5139)
    else:
        e.set_hbond_constraints(np.concatenate([ids, np.full((len(ids), 2), -1, np.int32)], 1),
                                np.tile(np.array([[0.9572, 1.0, 1.0]], np.float32), (len(ids), 1)))
    e.step(dt, 40)                                        # the first steps project the random velocities onto the rigid motion
    v = e.velocities().astype(np.float64)
    p_bar, vir = e.pressure()
    e.close()
    m = 1.0 / v[:, 3]
    ke = 0.5 * (m * (v[:, :3] ** 2).sum(1)).sum() / 418.4
    mm = m.reshape(-1, per)
    vcom = (mm[:, :, None] * v[:, :3].reshape(-1, per, 3)).sum(1) / mm.sum(1)[:, None]
    ke_com = 0.5 * (mm.sum(1) * (vcom ** 2).sum(1)).sum() / 418.4
    assert ke_com < 0.75 * ke                             # there is rotational energy to take out
    assert abs((2 * ke + vir) - 2 * ke_com) < 0.01 * 2 * ke, (ke, ke_com, vir)
    vol = float(np.prod(np.asarray(w["box_ext"], np.float64)))
    assert abs(p_bar - (2 * ke + vir) / (3 * vol) * 69476.95) < 1e-6 * abs(p_bar)


@pytest.mark.parametrize("direction", [+1, -1])
def test_berendsen_barostat_relaxes_the_box_towards_the_target(direction, Engine):
    """mc_set_barostat (BarostatCfg{pressure_target, tau}, reference ui/panels/md.rs:517-556): the volume moves the right way,
    the pressure ends near the target, the first scaling is the weak-coupling formula."""
    w = W.lj_fluid(m=12)
    e = Engine.from_workload(w)
    e.step(w["dt"], 10)
    p_start, _ = e.pressure()
    v_start = float(np.prod(e.box()[1] - e.box()[0]))
    e.close()
    target = p_start + (3000.0 if direction > 0 else -1500.0)
    beta, tau, every = 1e-4, 0.5, 10
    e = Engine.from_workload(w)
    e.set_barostat(1, target, tau_ps=tau, compressibility_per_bar=beta, every=every)
    e.step(w["dt"], 10)                                  # exactly one application, at the pressure measured above
    v1 = float(np.prod(e.box()[1] - e.box()[0]))
    mu3 = 1.0 - beta * (every * w["dt"] / tau) * (target - p_start)
    assert abs((v1 / v_start - 1.0) - (mu3 - 1.0)) < 0.05 * abs(mu3 - 1.0), (v1 / v_start, mu3)
    x = e.positions()
    lo, hi = e.box()
    assert np.all(x[:, :3] >= lo - 1e-3) and np.all(x[:, :3] <= hi + 1e-3)      # still inside the (scaled) box
    vols, ps = [], []
    for _ in range(8):
        e.step(w["dt"], 50)
        vols.append(float(np.prod(e.box()[1] - e.box()[0])))
        ps.append(e.pressure()[0])
    st = e.stats()
    e.close()
    assert (vols[-1] - v_start) * direction < 0           # compressed for a higher target, expanded for a lower one
    assert abs(ps[-1] - target) < 0.25 * abs(target - p_start), (ps, target)
    assert st["n_rebuilds"] >= 40                         # every application rebuilds the list for the new box


def test_npt_of_rigid_water_with_stochastic_cell_rescaling(Engine):
    """The reference's production set-up in one handle: rigid water (SETTLE), CSVR thermostat, barostat.  Stochastic cell
    rescaling needs the thermostat's temperature, is reproducible for a seed, keeps the molecules rigid and moves the box."""
    from molchanica_b200.engine import McError
    w = dict(W.water_box_c1(), coul_mode=2, alpha=0.35, skin=0.3)
    n = len(w["xyzq"])
    tri = np.arange(n, dtype=np.int32).reshape(-1, 3)

    def run(seed):
        e = Engine.from_workload(w)
        e.set_rigid_waters(tri, 0.9572, 1.5139)
        with pytest.raises(McError, match="thermostat"):
            e.set_barostat(2, 1.0, tau_ps=1.0, every=5, seed=seed)
        e.set_thermostat(2, 300.0, 10.0, seed=5)
        e.set_barostat(2, 1.0, tau_ps=1.0, compressibility_per_bar=4.5e-5, every=5, seed=seed)
        e.step(0.001, 60)
        x, box = e.positions(), e.box()
        e.close()
        return x, box
    x1, b1 = run(9)
    x2, b2 = run(9)
    x3, b3 = run(10)
    assert np.array_equal(b1[1], b2[1]) and np.allclose(x1, x2, atol=1e-4)
    assert not np.array_equal(b1[1], b3[1])               # another seed, another volume path
    ext = (b1[1] - b1[0]).astype(np.float64)
    assert np.all(np.abs(ext / np.asarray(w["box_ext"], np.float64) - 1.0) < 0.05) and abs(ext[0] - w["box_ext"][0]) > 1e-5
    m = x1[:, :3].astype(np.float64).reshape(-1, 3, 3)
    d = lambda a, b: np.linalg.norm((a - b) - np.rint((a - b) / ext) * ext, axis=1)
    # coordinates were scaled (bond lengths with them) at most once since the last SETTLE: rigid to 1e-3 A
    assert np.abs(d(m[:, 0], m[:, 1]) - 0.9572).max() < 2e-3 and np.abs(d(m[:, 1], m[:, 2]) - 1.5139).max() < 3e-3


def test_zero_com_drift_removes_the_net_momentum_of_the_mobile_atoms(Engine):
    """MdConfig.zero_com_drift (reference properties/crystal.rs:310): option zero_com_drift = k."""
    w = W.lj_fluid(m=12)
    w["vel"] = w["vel"].copy()
    w["vel"][:, 0] += 1.5                                   # the whole fluid drifts along x
    flags = np.zeros(len(w["xyzq"]), np.uint8)
    flags[::50] = 1                                         # static atoms: neither counted nor touched
    w["flags"] = flags
    mom = lambda v: ((v[:, :3] / v[:, 3:4]).astype(np.float64))[flags == 0].sum(0)
    e = Engine.from_workload(w)
    e.step(w["dt"], 4)
    p_free = mom(e.velocities())
    e.close()
    e = Engine.from_workload(w)
    e.set_option("zero_com_drift", 2)
    e.step(w["dt"], 4)
    v = e.velocities()
    e.close()
    assert abs(p_free[0]) > 1e4                            # without the option the drift stays
    assert np.abs(mom(v)).max() < 2e-4 * abs(p_free[0])    # with it the mobile atoms are at rest as a whole (forces between
    #                                                        mobile and static atoms feed a little back within two steps)
    assert np.array_equal(v[flags == 1], w["vel"][flags == 1])


def test_pressure_refuses_what_it_cannot_do(Engine):
    from molchanica_b200.engine import McError
    w = W.globule(200, seed=3)                          # vacuum: no volume
    e = Engine.from_workload(w)
    with pytest.raises(McError, match="periodic"):
        e.pressure()
    e.close()
    w = W.water_box_c1()
    e = Engine.from_workload(w)
    n = len(w["xyzq"])
    e.set_rigid_waters(np.arange(n, dtype=np.int32).reshape(-1, 3), 0.9572, 1.5139)
    with pytest.raises(McError, match="take a step first"):   # the constraint virial is that of the last step
        e.pressure()
    e.close()


def test_nve_energy_and_momentum_on_lj_fluid(Engine):
    w = W.lj_fluid(m=16)
    e = Engine.from_workload(w)
    e.compute_forces()
    en0 = e.energy()
    e.step(w["dt"], 200)
    en1 = e.energy()
    t0 = en0["energy_potential"] + en0["energy_kinetic"]
    t1 = en1["energy_potential"] + en1["energy_kinetic"]
    assert abs(t1 - t0) < 0.01 * en0["energy_kinetic"], (t0, t1)
    v = e.velocities()
    p = (v[:, :3] / v[:, 3:4]).astype(np.float64).sum(0)
    assert np.abs(p).max() < 1e-2 * float(np.abs(v[:, :3] / v[:, 3:4]).sum(0).max())
    assert e.stats()["n_rebuilds"] >= 2
    e.close()


def test_golden_fixtures(Engine):
    g = np.load(GOLD)
    for name in ("lj512", "water648", "glob300"):
        sc = g[f"{name}.scalars"]
        w = {k: g[f"{name}.{k}"] for k in ("xyzq", "vel", "type", "ljtab", "box_lo", "box_ext", "excl_start", "excl_idx", "pairs14")}
        w.update(periodic=bool(sc[0]), rc_lj=float(sc[1]), rc_q=float(sc[2]), skin=float(sc[3]), coul_mode=int(sc[4]),
                 alpha=float(sc[5]), scale14_lj=float(sc[6]), scale14_q=float(sc[7]), dt=float(sc[8]))
        e = Engine.from_workload(w)
        e.build_neighbors()
        start, idx = e.neighbors()
        assert np.array_equal(start, g[f"{name}.nbr_start"]) and np.array_equal(idx, g[f"{name}.nbr_idx"]), name
        e.compute_forces()
        assert force_rel_err(e.forces(), g[f"{name}.f64"], g[f"{name}.sumabs"]).max() < FORCE_RTOL, name
        assert energy_close(e.energy()["energy_potential_nonbonded"], g[f"{name}.energy"].sum(), g[f"{name}.f64"][:, 3]), name
        e.step(w["dt"], 10)
        ok, worst, scale = trajectory_close(e.positions(), g[f"{name}.x10"], w["xyzq"], w["box_ext"] if w["periodic"] else None)
        assert ok, (name, worst, scale)
        e.close()


def test_full_size_c4_properties(Engine):
    """BASELINE config 4 at full size (1,000,000 atoms): size-independent properties -- sampled rows
    bit-exact against a numpy brute force, list symmetry, Newton's third law, row-sum energy."""
    w = W.lj_fluid(m=100)
    n = len(w["xyzq"])
    e = Engine.from_workload(w)
    e.build_neighbors()
    start, idx = e.neighbors()
    assert abs(len(idx) / n - 80.0) < 3.0  # jittered simple-cubic lattice: ~80 entries per atom
    rng = np.random.default_rng(1)
    r_list = np.float32(w["rc_lj"]) + np.float32(w["skin"])
    for i in rng.integers(0, n, 24):
        assert np.array_equal(idx[start[i]:start[i + 1]], numpy_row(w["xyzq"], int(i), w["box_ext"], True, r_list)), i
    # symmetry on a sample of entries
    for i in rng.integers(0, n, 200):
        for j in idx[start[i]:start[i + 1]][:4]:
            row_j = idx[start[j]:start[j + 1]]
            assert row_j[np.searchsorted(row_j, i)] == i
    e.compute_forces()
    f = e.forces().astype(np.float64)
    assert np.abs(f[:, :3].sum(0)).max() < 1e-6 * np.abs(f[:, :3]).sum()
    assert abs(e.energy()["energy_potential_nonbonded"] - 0.5 * f[:, 3].sum()) < 1e-6 * abs(f[:, 3].sum())
    e.close()


@pytest.mark.parametrize("uniform", [0, 1])
def test_pair_loop_variants_agree_with_the_oracle(uniform, Engine, oracle):
    """Option pair_uniform switches to the warp-uniform row loop (warp-wide skin-shell skipping)."""
    for w in (W.solvated_c3(), W.lj_fluid(m=20)):
        e = Engine.from_workload(w)
        e.set_option("pair_uniform", uniform)
        e.compute_forces()
        nb = oracle.neighbors(w)
        f64, scale, en = oracle.forces(w, nb, precision=64)
        assert force_rel_err(e.forces(), f64, scale).max() < FORCE_RTOL
        assert energy_close(e.energy()["energy_potential_nonbonded"], en.sum(), f64[:, 3])
        e.close()


def test_blow_up_is_reported_not_crashed(Engine):
    """Two overlapping atoms and an absurd time step: the engine must return an error (non-finite
    coordinates) and leave the device usable -- the reference panics on CUDA errors
    (src/reflection.rs:200); a library must not."""
    from molchanica_b200.engine import McError
    w = W.lj_fluid(m=8)
    w["xyzq"][1, :3] = w["xyzq"][0, :3] + np.float32(1e-4)
    e = Engine.from_workload(w)
    with pytest.raises(McError):
        for _ in range(50):
            e.step(0.5, 10)
    e.close()
    e2 = Engine.from_workload(W.lj_fluid(m=8))  # the context survived
    e2.step(0.002, 5)
    assert np.isfinite(e2.positions()).all()
    e2.close()
