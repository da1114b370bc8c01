"""MD paths beside the plain NVE step, through the C ABI on the GPU: the pipelined external-force upload, pressure /
virial, constraint virial, barostats, drift removal, snapshots with velocities.  First run on hardware by the driver at
the end of round 1 (all passed); tests/test_library_on_host.py also runs this file against the host build of the library."""
import numpy as np
import pytest

from molchanica_b200 import workloads as W
from util import trajectory_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Engine():
    from molchanica_b200.engine import MdEngine
    return MdEngine


def test_snapshot_with_velocities(Engine):
    """mc_snapshot_begin_pv (Snapshot.atom_velocities, reference src/md/trajectory.rs:160-204), also right after a pipelined
    call with external forces."""
    w = W.lj_fluid(m=12)
    n = len(w["xyzq"])
    e = Engine.from_workload(w)
    ext = np.zeros((n, 3), np.float32)
    ext[::5, 1] = 2.0
    e.step(w["dt"], 2, ext_forces=ext)
    px, pv = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
    assert e.snapshot_begin_pv(px, pv) == n
    e.snapshot_wait()
    assert np.array_equal(px, e.positions()) and np.array_equal(pv, e.velocities())
    e.close()


def _run_with_changing_ext(Engine, w, defer, n_calls, steps_per_call, poke):
    """One mc_step(dt, k, ext) per call with a different array every call, positions read back after every call (as the
    reference's alignment loop does, src/mol_alignment.rs:318-353).  `poke` = observers / setters thrown in on the way."""
    n = len(w["xyzq"])
    e = Engine.from_workload(w)
    e.set_option("fused_steps", 0)  # the per-launch path with and without the deferral is what is compared bit for bit here
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


@pytest.mark.parametrize("case", ["globule", "bonded_globule", "hot_fluid", "cold_fluid", "odd_sizes"])
def test_fused_multi_step_kernel_equals_the_per_launch_path(case, Engine, oracle):
    """Small plain-NVE systems take all steps of a call in one cooperative launch (md_fused.cu, options fused_steps / fused_brute, the
    GUI's ten-steps-per-frame path of reference src/md/mod.rs:45,737-749).  Same arithmetic as the per-launch path: with no
    rebuild inside the run and the same lanes per row positions, velocities and forces are bit-identical; with rebuilds (the fused kernel uses the
    synchronous displacement criterion, the per-launch path the look-ahead one: lists are rebuilt at different steps and rows
    change their summation order) both follow the oracle's trajectory.  System sizes that leave the last warp of the row
    loop partly empty are part of the sweep (the first hardware run of this kernel hung on exactly that)."""
    if case == "globule":
        ws, bonded, n_calls, k = [W.globule()], False, 3, 2  # (atoms without bonds: a few steps, as in test_gpu_parity.py)
    elif case == "bonded_globule":
        ws, bonded, n_calls, k = [W.bonded_globule()], True, 3, 4
    elif case == "hot_fluid":
        ws, bonded, n_calls, k = [dict(W.lj_fluid(m=12, temp_k=400.0), skin=0.6)], False, 4, 15
    elif case == "cold_fluid":  # no rebuild inside the run: the host's rows inside the fused kernel give the per-launch path's bits
        ws, bonded, n_calls, k = [W.lj_fluid(m=10)], False, 3, 10
    else:
        ws, bonded, n_calls, k = [W.globule(n, seed=300 + n) for n in (1, 2, 3, 5, 31, 33, 127, 257)], False, 2, 3
    for w in ws:
        n = len(w["xyzq"])
        runs = []
        for fused, brute in ((1, 0), (0, 0), (1, 1)):  # the host's rows inside the fused kernel / per launch / the kernel's own list
            e = Engine.from_workload(w, bonded=bonded)
            e.set_option("fused_steps", fused)
            e.set_option("fused_brute", brute)
            if case == "cold_fluid":
                e.set_option("fused_lanes", 8)  # the lane count of the per-launch pair kernel: same summation order
            ext = np.zeros((n, 3), np.float32)
            ext[::3, 0] = 1.5
            for c in range(n_calls):
                e.step(w["dt"], k, ext_forces=ext if c == 1 else None)
            runs.append(dict(x=e.positions(), v=e.velocities(), f=e.forces(), st=e.stats()))
            e.close()
        a, b, c = runs
        assert a["st"]["n_steps"] == b["st"]["n_steps"] == c["st"]["n_steps"] == n_calls * k
        # one launch per call (+ what a rebuild inside the call adds) against two and more per step
        import os
        if not os.environ.get("MOLCHANICA_MD_LIB"):  # (the host build of the library has no cooperative launch: both runs are per-launch there)
            assert a["st"]["n_kernel_launches"] <= b["st"]["n_kernel_launches"] - n_calls * (k - 1)  # (list builds are in both counts)
            assert c["st"]["n_kernel_launches"] <= n_calls + 4  # one launch per call (+ the reads of positions, velocities, forces), no list build
        if case == "cold_fluid" and not os.environ.get("MOLCHANICA_MD_LIB"):
            assert a["st"]["n_rebuilds"] == 1 and b["st"]["n_rebuilds"] == 1
            assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["v"], b["v"]) and np.array_equal(a["f"][:, :3], b["f"][:, :3])
        if case == "hot_fluid":
            assert a["st"]["n_rebuilds"] >= 3
        if not bonded and case != "odd_sizes":
            cur = dict(xyzq=w["xyzq"], vel=w["vel"])  # replay the calls on the oracle (the second one carries external forces)
            for call in range(n_calls):
                cur = oracle.md_run(w, k, precision=64, xyzq=cur["xyzq"], vel=cur["vel"], ext_force=ext if call == 1 else None)
            for r in (a, b, c):
                ok, worst, scale = trajectory_close(r["x"], cur["xyzq"], w["xyzq"], w["box_ext"] if w["periodic"] else None)
                assert ok, (case, worst, scale)
        else:
            for r in (b, c):
                ok, worst, scale = trajectory_close(a["x"], r["x"], w["xyzq"], w["box_ext"] if w["periodic"] else None)
                assert ok, (case, n, worst, scale)
        if case == "hot_fluid" and not os.environ.get("MOLCHANICA_MD_LIB"):
            assert c["st"]["n_rebuilds"] >= 3  # rebuilt inside the launches


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


def test_minimiser_keeps_virtual_sites_on_their_parents(Engine):
    """mc_minimize_energy moves atoms with the step path's drift, which leaves the massless site M of a four-site water where
    it was: the minimiser has to place M again after every move (M = O + a (H1 - O) + b (H2 - O)), or sites and charges go
    stale.  Flexible OPC waters (no constraints: the minimiser refuses those), 15 iterations."""
    w = W.water_box_opc()
    e = Engine.from_workload(w)
    a, b = w["vsite_ab"]
    e.set_virtual_sites(w["virtual_sites"], a, b)
    x0 = e.positions()
    acc, e0, e1 = e.minimize_energy(15)
    x = e.positions()[:, :3].astype(np.float64)
    assert acc >= 1 and e1 <= e0 and np.abs(x - x0[:, :3]).max() > 1e-4      # something moved, downhill
    q = np.asarray(w["virtual_sites"])
    ext = np.asarray(w["box_ext"], np.float64)
    mi = lambda d: d - np.rint(d / ext) * ext
    want = x[q[:, 1]] + a * mi(x[q[:, 2]] - x[q[:, 1]]) + b * mi(x[q[:, 3]] - x[q[:, 1]])
    assert np.abs(mi(x[q[:, 0]] - want)).max() < 2e-5
    e.close()


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


def test_cpp_host_mirror_npt_configuration():
    """include/molchanica_md.hpp configured the way properties/crystal.rs:306-316 configures `dynamics` (thermostat, drift
    removal, barostat), pressure in the energy snapshot, snapshots with velocities: tests/cpp/host_mirror_smoke.cpp --npt
    against whichever library this process uses."""
    import os
    import subprocess
    import tempfile

    import __graft_entry__ as g
    from molchanica_b200 import _lib
    g.build_cpp_host()
    here = os.path.dirname(os.path.abspath(__file__))
    d = tempfile.mkdtemp()
    os.symlink(_lib.LIB_PATH, os.path.join(d, "libmolchanica_md.so"))
    env = dict(os.environ, LD_LIBRARY_PATH=d + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([os.path.join(here, "cpp", "_build", "host_mirror_smoke"), "--npt"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "host mirror ok" in r.stdout, r.stdout + r.stderr
