"""Parity of the CUDA path (through the C ABI) with the oracle: neighbour indices bit-exact,
forces within 1e-5 of the fp64 truth relative to sum|f_ij|, energies within 1e-5, short
trajectories, and the committed golden fixtures.  Needs a B200."""
import os

import numpy as np
import pytest

from molchanica_b200 import workloads as W
from util import FORCE_RTOL, FORCE_RTOL_NET, energy_close, force_rel_err, numpy_row, trajectory_close

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
    _, sumnet, _ = oracle.forces(w, (o_start, o_idx), precision=64, scale="net")
    err_net = force_rel_err(f, f64, sumnet)
    assert err_net.max() < FORCE_RTOL_NET, f"max force error on the net-pair scale {err_net.max():.3e} at atom {err_net.argmax()}"
    assert energy_close(e.energy()["energy_potential_nonbonded"], en.sum(), f64[:, 3])
    # per-atom energy rows
    assert np.abs(f[:, 3] - f64[:, 3]).max() < 1e-5 * max(1.0, float(np.abs(f64[:, 3]).max()))
    e.close()


@pytest.mark.parametrize("name", ["lj1728", "lj8000", "ions_plain", "ions_erfc", "ions_nocoul", "water648"])
def test_tile_kernel_matches_the_oracle(name, Engine, oracle):
    """pair_tile.cu (TMA-staged tile + row block, compact rows of 16-bit tile-local indices, packed fp32) forced on for
    systems below its automatic size threshold: one and many LJ types, no / plain / erfc Coulomb, energies, wrapping and
    interior cells, and the expansion of the compact rows back to global indices (mc_get_neighbors) -- same bars as above.
    water648 is the counter-example: one cell holds all 648 atoms, the build must notice and fall back to global-slot rows."""
    w = {"ions_plain": lambda: W.ionic_mixture(12, coul_mode=1), "ions_erfc": lambda: W.ionic_mixture(12, coul_mode=2),
         "ions_nocoul": lambda: W.ionic_mixture(12, coul_mode=0)}.get(name, _cases().get(name))()
    e = Engine.from_workload(w)
    e.set_option("pair_tile", 1)
    e.build_neighbors()
    st = e.stats()
    compact = st["list_bytes"] < 2.2 * max(st["n_pairs_listed"], 1) + 64 * len(w["xyzq"])
    assert compact == (name != "water648"), "compact rows in use: %s" % compact
    start, idx = e.neighbors()
    o_start, o_idx = oracle.neighbors(w)
    assert np.array_equal(start, o_start) and np.array_equal(idx, o_idx)
    e.compute_forces()
    f = e.forces()
    f64, sumabs, en = oracle.forces(w, (o_start, o_idx), precision=64)
    err = force_rel_err(f, f64, sumabs)
    assert err.max() < FORCE_RTOL, f"max force error {err.max():.3e} at atom {err.argmax()}"
    assert energy_close(e.energy()["energy_potential_nonbonded"], en.sum(), f64[:, 3])
    assert np.abs(f[:, 3] - f64[:, 3]).max() < 1e-5 * max(1.0, float(np.abs(f64[:, 3]).max()))
    n_steps = 20 if name.startswith("lj") else 6
    dt = w["dt"]
    e.step(dt, n_steps)   # the step path runs the instantiation without energies (packed arithmetic for LJ-only systems)
    ref = oracle.md_run(w, n_steps, precision=64)
    ok, worst, scale = trajectory_close(e.positions(), ref["xyzq"], w["xyzq"], w["box_ext"] if w["periodic"] else None)
    assert ok, (worst, scale)
    # forces of the packed / energy-free instantiation against the gather kernel on the positions just reached
    e.compute_forces()
    f_tile = e.forces()
    e.set_option("pair_tile", 0)
    e.compute_forces()
    f_gather = e.forces()
    scale = np.abs(f_gather[:, :3]).max() + 1e-30
    assert np.abs(f_tile[:, :3] - f_gather[:, :3]).max() < 2e-5 * scale
    e.close()


@pytest.mark.parametrize("name", ["lj1728", "water648", "solv23558"])
def test_both_list_build_kernels_write_the_same_rows(name, Engine, oracle):
    """rows_build_kernel (default: ballot compaction, rows staged in shared memory, decoupled warps) and tile_build_kernel
    (option build_variant = 1) list the oracle's pairs bit for bit -- on the first build (no row-length hint yet: every row
    takes the two-sweep path), on a rebuild (rows staged), and with the staging space capped below the row length (mixed) --
    and the forces computed from either list are the same numbers (same row order: ascending tile index).
    Dense systems (solv23558: a 148 KB tile, one CTA per SM) take rows_build_kernel's other configuration: 16 consumer warps,
    first build count + second sweep, later builds ONE sweep into space claimed from the previous build's longest row; with
    the hint capped (limit 1) that space overflows and the build must fall back; rows_dense = 0 is round 2's configuration."""
    w = _cases()[name]()
    o_start, o_idx = oracle.neighbors(w)
    forces = []
    for variant, limit, dense in ((1, 0, 1), (2, 0, 1), (2, 1, 1), (2, 0, 0)):
        e = Engine.from_workload(w)
        e.set_option("build_variant", variant)
        e.set_option("row_stage_limit", limit)
        e.set_option("rows_dense", dense)
        for rep in range(3):  # later builds: the hint of the first one is there
            e.build_neighbors()
            start, idx = e.neighbors()
            assert np.array_equal(start, o_start) and np.array_equal(idx, o_idx), (variant, limit, dense, rep)
            e.set_positions(w["xyzq"])  # invalidates the list
        e.compute_forces()
        forces.append(e.forces())
        e.close()
    assert all(np.array_equal(forces[0], f) for f in forces[1:])


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


@pytest.mark.parametrize("name", ["lj1728", "lj8000", "water648", "glob1231", "solv23558"])
def test_interleaved_rows_are_the_same_rows(name, Engine):
    """The 8-lane force kernel can read a quad-interleaved copy of the rows (option rows_interleave, default off: the chunks
    of 8 entries of four consecutive rows alternate, one 128-byte line per warp-wide index load).  Same entries, same lanes, same
    order of accumulation: forces and per-atom energies are the same BITS as from the plain rows, with and without energies,
    and after steps through rebuilds."""
    w = _cases()[name]()
    out = []
    for ilv in (1, 0):
        e = Engine.from_workload(w)
        e.set_option("rows_interleave", ilv)
        e.compute_forces()
        f0 = e.forces()
        e.step(w["dt"], 12)
        x = e.positions()
        e.compute_forces()
        out.append((f0, x, e.forces(), e.stats()["n_rebuilds"]))
        e.close()
    for a, b in zip(out[0][:3], out[1][:3]):
        assert np.array_equal(a, b)
    assert out[0][3] == out[1][3]


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


def test_packed_xyz_snapshot(Engine):
    """mc_snapshot_begin_xyz: the same hand-off as mc_snapshot_begin, 3 floats per atom in the caller's order."""
    w = W.lj_fluid(m=12)
    e = Engine.from_workload(w)
    e.step(w["dt"], 7)
    xyz = np.zeros((len(w["xyzq"]), 3), np.float32)
    n, epoch = e.snapshot_begin_xyz(xyz)
    e.snapshot_wait()
    assert n == len(w["xyzq"]) and epoch == e.stats()["n_rebuilds"]
    assert np.array_equal(xyz, e.positions()[:, :3])
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


def test_full_size_c4_matches_the_oracle(Engine, oracle):
    """BASELINE config 4 at full size against the oracle itself, ALL atoms: the 80,000,000-entry Verlet list bit-exact
    with the oracle's cell-list build, forces of every atom within 1e-5 of the fp64 sum over that list (and 3e-5 on the
    net-pair scale; see the two truths below), the energy, and a 20-step trajectory against the oracle's fp64 path."""
    w = W.lj_fluid(m=100)
    n = len(w["xyzq"])
    e = Engine.from_workload(w)
    e.build_neighbors()
    start, idx = e.neighbors()
    o_start, o_idx = oracle.neighbors(w)
    assert np.array_equal(start, o_start), "row lengths differ"
    assert np.array_equal(idx, o_idx), "neighbour indices differ"
    del start, idx
    e.compute_forces()
    f = e.forces()
    # (1) against fp64 arithmetic on the reference's own inputs -- the fp32 minimum-image difference it forms
    #     (float3 diff = posit_tgt - posit_src, util.cu:65-71): every atom within 1e-5
    f64r, sumabs, en = oracle.forces(w, (o_start, o_idx), precision=6432)
    err = force_rel_err(f, f64r, sumabs)
    assert err.max() < FORCE_RTOL, f"max force error {err.max():.3e} at atom {err.argmax()} of {n}"
    _, sumnet, _ = oracle.forces(w, (o_start, o_idx), precision=6432, scale="net")
    err_net = force_rel_err(f, f64r, sumnet)
    assert err_net.max() < FORCE_RTOL_NET, f"net-pair scale: {err_net.max():.3e} at atom {err_net.argmax()}"
    # (2) against the exact fp64 differences.  Atoms farther than the list radius from every box face: 1e-5 as always.
    #     Atoms at the periodic seam: a difference taken across it (x_i ~ 360, x_j ~ 0) is rounded to ulp(360) = 3e-5 A
    #     BEFORE the minimum image, in the reference (fp32 positions, fp32 subtraction) exactly as here -- the oracle's own
    #     reference-form fp32 path has its worst atom (1.30e-5, atom 400078) at the same place -- so those are held to 3e-5.
    f64, sumabs_x, _ = oracle.forces(w, (o_start, o_idx), precision=64)
    err_x = force_rel_err(f, f64, sumabs_x)
    x, L = w["xyzq"][:, :3], np.asarray(w["box_ext"], np.float32)
    r_list = w["rc_lj"] + w["skin"]
    seam = ((x < r_list + 0.1) | (x > L - r_list - 0.1)).any(1)
    assert err_x[~seam].max() < FORCE_RTOL, f"interior atoms: {err_x[~seam].max():.3e}"
    assert err_x[seam].max() < 3e-5, f"atoms at the periodic seam: {err_x[seam].max():.3e}"
    f32r, _, _ = oracle.forces(w, (o_start, o_idx), precision=32)
    assert force_rel_err(f32r, f64, sumabs_x)[seam].max() > 0.5 * err_x[seam].max()  # the reference-form arithmetic is no better there
    assert energy_close(e.energy()["energy_potential_nonbonded"], en.sum(), f64[:, 3])
    del o_start, o_idx
    e.step(w["dt"], 20)
    ref = oracle.md_run(w, 20, precision=64)
    ok, worst, scale = trajectory_close(e.positions(), ref["xyzq"], w["xyzq"], w["box_ext"])
    assert ok, (worst, scale)
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
