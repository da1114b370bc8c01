"""Synthetic, seeded stand-ins for the five BASELINE.json configurations.

The reference ships no Amber parameter files or structures (reference .gitignore:15-34), so
every configuration is generated here from fixed seeds.  A workload is a plain dict of numpy
arrays in the layout the C ABI ingests (include/molchanica_md.h):

    xyzq      (n,4) f32   x, y, z [A], charge pre-scaled by sqrt(332.0522)  
    vel       (n,4) f32   vx, vy, vz [A/ps], inverse mass [1/amu] (0 = static atom)
    type      (n,)  u16   LJ type id
    ljtab     (T,T,2) f32 (sigma_ij [A], eps_ij [kcal/mol]), Lorentz-Berthelot from Amber Rmin/2
    box_lo, box_ext (3,) f32; periodic (bool)
    rc_lj, rc_q, skin [A]; coul_mode (0 none, 1 plain cutoff, 2 erfc real-space); alpha
    excl_start (n+1,) i32, excl_idx i32   CSR of excluded partners (1-2, 1-3, 1-4)
    pairs14   (m,2) i32, scale14_lj, scale14_q
    dt [ps]
"""
from __future__ import annotations

import numpy as np

COULOMB_SCALE = float(np.sqrt(332.0522))  # charges carry sqrt(k_e) so q_i*q_j/r is kcal/mol
KB = 0.0019872041  # kcal/mol/K
ACCEL_CONV = 418.4  # kcal/mol/A/amu -> A/ps^2

# Amber-like LJ classes: Rmin/2 [A], eps [kcal/mol], mass [amu]
_CLASSES = {
    "H": (1.20, 0.0157, 1.008),
    "C": (1.908, 0.1094, 12.011),
    "N": (1.824, 0.1700, 14.007),
    "O": (1.6612, 0.2100, 15.999),
    "S": (2.000, 0.2500, 32.06),
    "OW": (1.7683, 0.1521, 15.999),  # TIP3P oxygen: sigma 3.15061 A
    "HW": (0.5, 0.0, 1.008),  # TIP3P hydrogen: no LJ
    "AR": (3.405 * 2 ** (1 / 6) / 2, 0.2381, 39.948),
}


def lj_table(names):
    """Lorentz-Berthelot (sigma_ij, eps_ij) table from Amber Rmin/2, eps."""
    T = len(names)
    tab = np.zeros((T, T, 2), np.float32)
    for a, na in enumerate(names):
        for b, nb in enumerate(names):
            ra, ea, _ = _CLASSES[na]
            rb, eb, _ = _CLASSES[nb]
            tab[a, b, 0] = (ra + rb) * 2 ** (-1 / 6)
            tab[a, b, 1] = np.sqrt(ea * eb)
    return tab


def _empty_topology(n):
    return np.zeros(n + 1, np.int32), np.zeros(0, np.int32), np.zeros((0, 2), np.int32)


def topology_from_bonds(n, bonds):
    """1-2/1-3/1-4 exclusion CSR and the 1-4 pair list from a bond graph (what
    MdState::new derives from `MolDynamics.bonds`, reference src/md/mod.rs:1110-1151)."""
    adj = [[] for _ in range(n)]
    for i, j in bonds:
        adj[i].append(j)
        adj[j].append(i)
    excl = [set() for _ in range(n)]
    pairs14 = set()
    for i in range(n):
        d1 = set(adj[i])
        d2 = set()
        for j in d1:
            d2.update(adj[j])
        d2 -= d1 | {i}
        d3 = set()
        for j in d2:
            d3.update(adj[j])
        d3 -= d1 | d2 | {i}
        excl[i] = d1 | d2 | d3
        for j in d3:
            if i < j:
                pairs14.add((i, j))
    start = np.zeros(n + 1, np.int32)
    idx = []
    for i in range(n):
        row = sorted(excl[i])
        idx.extend(row)
        start[i + 1] = len(idx)
    p14 = np.array(sorted(pairs14), np.int32).reshape(-1, 2)
    return start, np.array(idx, np.int32), p14


def _saw_globule(n_atoms, seed, heavy_frac=0.5, density=0.05):
    """Protein-like globule, built without any unbounded search: heavy atoms by random
    sequential addition inside a sphere (min separation 2.6 A), chained into one molecule by a
    greedy nearest-neighbour path (the bond graph only feeds the exclusion / 1-4 topology; there
    are no bonded forces on this path), then hydrogens at 1.0 A from round-robin parents.
    Returns positions (n,3) f64, class names, bonds."""
    rng = np.random.default_rng(seed)
    n_heavy = max(1, int(round(n_atoms * heavy_frac)))
    n_h = n_atoms - n_heavy
    radius = max((3.0 * n_atoms / (4.0 * np.pi * density)) ** (1 / 3),
                 (n_heavy * 9.2 / 0.22 * 3.0 / (4.0 * np.pi)) ** (1 / 3))
    pos = np.zeros((n_atoms, 3))
    k, tries = 0, 0
    while k < n_heavy:
        cand = rng.uniform(-radius, radius, size=(256, 3))
        cand = cand[np.linalg.norm(cand, axis=1) <= radius]
        for c in cand:
            if k == n_heavy:
                break
            if k == 0 or np.min(np.linalg.norm(pos[:k] - c, axis=1)) >= 2.6:
                pos[k] = c
                k += 1
        tries += 1
        if tries % 200 == 0:
            radius *= 1.05  # jammed: give it room (keeps the loop bounded)
    # greedy nearest-neighbour chain over the heavy atoms
    order = [0]
    left = np.ones(n_heavy, bool)
    left[0] = False
    for _ in range(n_heavy - 1):
        d = np.linalg.norm(pos[:n_heavy] - pos[order[-1]], axis=1)
        d[~left] = np.inf
        j = int(np.argmin(d))
        order.append(j)
        left[j] = False
    pos[:n_heavy] = pos[order]
    bonds = [(i - 1, i) for i in range(1, n_heavy)]
    parents = rng.permutation(n_heavy)
    for h in range(n_h):
        k = n_heavy + h
        par = int(parents[h % n_heavy])
        best, best_d = None, -1.0
        for _ in range(24):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            cand = pos[par] + 1.0 * d
            dist = np.linalg.norm(pos[:k] - cand, axis=1)
            dist[par] = 9.0
            m = float(dist.min())
            if m > best_d:
                best, best_d = cand, m
            if m >= 1.9:
                break
        pos[k] = best
        bonds.append((par, k))
    p = np.array([32.0, 8.5, 9.0, 0.5])
    heavy_cls = rng.choice(["C", "N", "O", "S"], size=n_heavy, p=p / p.sum())
    names = list(heavy_cls) + ["H"] * n_h
    return pos, names, bonds


def _pack(pos, charges, masses, names, class_order, **kw):
    n = len(pos)
    xyzq = np.zeros((n, 4), np.float32)
    xyzq[:, :3] = pos
    xyzq[:, 3] = np.asarray(charges) * COULOMB_SCALE
    vel = np.zeros((n, 4), np.float32)
    vel[:, 3] = 1.0 / np.asarray(masses)
    lut = {c: t for t, c in enumerate(class_order)}
    typ = np.array([lut[c] for c in names], np.uint16)
    w = dict(xyzq=xyzq, vel=vel, type=typ, ljtab=lj_table(class_order), alpha=0.35,
             scale14_lj=0.5, scale14_q=1.0 / 1.2)
    w.update(kw)
    return w


def maxwell_boltzmann(vel, temp_k, seed):
    """Seeded Maxwell-Boltzmann velocities [A/ps] with centre-of-mass motion removed."""
    rng = np.random.default_rng(seed)
    n = len(vel)
    inv_m = vel[:, 3].astype(np.float64)
    sig = np.sqrt(KB * temp_k * inv_m * ACCEL_CONV)
    v = rng.normal(size=(n, 3)) * sig[:, None]
    m = np.where(inv_m > 0, 1.0 / np.maximum(inv_m, 1e-30), 0.0)
    v -= (m[:, None] * v).sum(0) / m.sum()
    vel[:, :3] = v.astype(np.float32)
    return vel


def _water_geometry(rng):
    """Three-site water in a random orientation, oxygen at the origin (TIP3P geometry)."""
    r_oh, ang = 0.9572, np.deg2rad(104.52)
    h0 = np.array([r_oh, 0, 0])
    h1 = np.array([r_oh * np.cos(ang), r_oh * np.sin(ang), 0])
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    rot = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                    [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                    [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return np.stack([np.zeros(3), rot @ h0, rot @ h1])


def water_box_c1(seed=101):
    """C1: 216 three-site waters (648 atoms), TIP3P constants, cubic PBC L = 18.64 A."""
    rng = np.random.default_rng(seed)
    L, m = 18.64, 6
    pos, names, charges, masses, bonds, kr0 = [], [], [], [], [], []
    for a in range(m):
        for b in range(m):
            for c in range(m):
                o = (np.array([a, b, c]) + 0.5) * (L / m)
                g = _water_geometry(rng) + o
                base = len(pos)
                pos.extend(g)
                names += ["OW", "HW", "HW"]
                charges += [-0.834, 0.417, 0.417]
                masses += [15.999, 1.008, 1.008]
                bonds += [(base, base + 1), (base, base + 2), (base + 1, base + 2)]
                kr0 += [(450.0, 0.9572), (450.0, 0.9572), (100.0, 1.5139)]
    n = len(pos)
    es, ei, _ = topology_from_bonds(n, bonds)
    w = _pack(np.array(pos), charges, masses, names, ["OW", "HW"],
              box_lo=np.zeros(3, np.float32), box_ext=np.full(3, L, np.float32), periodic=True,
              rc_lj=9.0, rc_q=9.0, skin=0.3, coul_mode=2, excl_start=es, excl_idx=ei,
              pairs14=np.zeros((0, 2), np.int32), dt=0.0005, name="C1-water216")
    w["bonds"] = np.array(bonds, np.int32)
    w["bond_kr0"] = np.array(kr0, np.float32)
    maxwell_boltzmann(w["vel"], 300.0, seed + 1)
    return w


OPC = dict(r_oh=0.8724, angle=np.deg2rad(103.6), d_om=0.1594, q_h=0.6791, sigma_o=3.16655, eps_o=0.21280)


def water_box_opc(seed=111, m=6, L=18.64):
    """m^3 four-site OPC waters (O, H, H, M per molecule; the reference's water model, ui/panels/md.rs:393): LJ on O,
    charges on H and on the massless site M = O + a [(H1 - O) + (H2 - O)]; rigid (SETTLE) + virtual site."""
    rng = np.random.default_rng(seed)
    r_oh, ang = OPC["r_oh"], OPC["angle"]
    a = OPC["d_om"] / (2.0 * r_oh * np.cos(ang / 2))
    pos, names, charges, masses, bonds = [], [], [], [], []
    for i in range(m):
        for j in range(m):
            for k in range(m):
                o = (np.array([i, j, k]) + 0.5) * (L / m)
                g = _water_geometry(rng)
                # rescale the TIP3P template to the OPC geometry: same plane, OPC bond length and angle
                h0 = g[1] / np.linalg.norm(g[1])
                perp = g[2] - (g[2] @ h0) * h0
                perp /= np.linalg.norm(perp)
                g1, g2 = r_oh * h0, r_oh * (np.cos(ang) * h0 + np.sin(ang) * perp)
                base = len(pos)
                pos.extend([o, o + g1, o + g2, o + a * (g1 + g2)])
                names += ["OPC_O", "HW", "HW", "HW"]
                charges += [0.0, OPC["q_h"], OPC["q_h"], -2 * OPC["q_h"]]
                masses += [15.999, 1.008, 1.008, np.inf]
                bonds += [(base, base + 1), (base, base + 2), (base + 1, base + 2), (base, base + 3), (base + 1, base + 3), (base + 2, base + 3)]
    n = len(pos)
    es, ei, _ = topology_from_bonds(n, bonds)
    _CLASSES["OPC_O"] = (OPC["sigma_o"] * 2 ** (1 / 6) / 2, OPC["eps_o"], 15.999)
    w = _pack(np.array(pos), charges, masses, names, ["OPC_O", "HW"],
              box_lo=np.zeros(3, np.float32), box_ext=np.full(3, L, np.float32), periodic=True,
              rc_lj=9.0, rc_q=9.0, skin=0.3, coul_mode=2, excl_start=es, excl_idx=ei,
              pairs14=np.zeros((0, 2), np.int32), dt=0.001, name=f"OPC-water{m ** 3}")
    flags = np.zeros(n, np.uint8)
    flags[3::4] = 1                                     # M: static to the integrator
    w["flags"] = flags
    idx = np.arange(n, dtype=np.int32).reshape(-1, 4)
    w["rigid_waters"] = np.ascontiguousarray(idx[:, :3])
    w["virtual_sites"] = np.ascontiguousarray(idx[:, [3, 0, 1, 2]])
    w["vsite_ab"] = (float(a), float(a))
    w["d_oh"], w["d_hh"] = float(r_oh), float(2 * r_oh * np.sin(ang / 2))
    maxwell_boltzmann(w["vel"], 300.0, seed + 1)
    w["vel"][3::4, :3] = 0.0
    return w


def globule(n_atoms=1231, seed=202, name="C2-globule1231", temp_k=300.0):
    """C2: protein-like globule in vacuum (non-periodic), Amber-like types, 1-2/1-3 exclusions,
    1-4 scaling, r_c 12 A, skin 2 A, dt 2 fs (reference default, src/prefs/mod.rs:203)."""
    pos, names, bonds = _saw_globule(n_atoms, seed)
    rng = np.random.default_rng(seed + 7)
    q = rng.normal(0.0, 0.35, n_atoms)
    q -= q.mean()
    masses = [_CLASSES[c][2] for c in names]
    es, ei, p14 = topology_from_bonds(n_atoms, bonds)
    lo = pos.min(0) - 1.0
    ext = pos.max(0) + 1.0 - lo
    w = _pack(pos, q, masses, names, ["H", "C", "N", "O", "S"],
              box_lo=lo.astype(np.float32), box_ext=ext.astype(np.float32), periodic=False,
              rc_lj=12.0, rc_q=12.0, skin=2.0, coul_mode=1, excl_start=es, excl_idx=ei,
              pairs14=p14, dt=0.002, name=name)
    maxwell_boltzmann(w["vel"], temp_k, seed + 1)
    return w


def bonded_terms_from_bonds(pos, bonds, seed=0):
    """Angles and dihedrals of a bond graph with Amber-like parameters: harmonic bonds and angles whose
    equilibrium values are the current geometry displaced by a few percent (so that forces are non-zero),
    periodic dihedrals with n in {1, 2, 3} and phase 0 or pi."""
    rng = np.random.default_rng(9000 + seed)
    n = len(pos)
    adj = [[] for _ in range(n)]
    for i, j in bonds:
        adj[i].append(j)
        adj[j].append(i)
    b = np.array(sorted({(min(i, j), max(i, j)) for i, j in bonds}), np.int32).reshape(-1, 2)
    r0 = np.linalg.norm(pos[b[:, 0]] - pos[b[:, 1]], axis=1)
    bk = np.stack([rng.uniform(200, 450, len(b)), r0 * rng.uniform(0.97, 1.03, len(b))], 1).astype(np.float32)
    ang = np.array(sorted({(min(i, k), j, max(i, k)) for j in range(n) for i in adj[j] for k in adj[j] if i < k}),
                   np.int32).reshape(-1, 3)
    va, vb = pos[ang[:, 0]] - pos[ang[:, 1]], pos[ang[:, 2]] - pos[ang[:, 1]]
    th = np.arccos(np.clip((va * vb).sum(1) / (np.linalg.norm(va, axis=1) * np.linalg.norm(vb, axis=1)), -1, 1))
    ak = np.stack([rng.uniform(30, 80, len(ang)), th + rng.uniform(-0.08, 0.08, len(ang))], 1).astype(np.float32)
    dih = set()
    for j, k in b:
        for i in adj[j]:
            for l in adj[k]:
                if i != k and l != j and i != l:
                    dih.add((i, int(j), int(k), l))
    dih = np.array(sorted(dih), np.int32).reshape(-1, 4)
    dk = np.stack([rng.uniform(0.1, 2.5, len(dih)), rng.integers(1, 4, len(dih)).astype(np.float64),
                   np.pi * rng.integers(0, 2, len(dih))], 1).astype(np.float32)
    return dict(bonds=b, bond_kr0=bk, angles=ang, angle_kt0=ak, dihedrals=dih, dihedral_prm=dk)


def hbond_clusters(names, bonds, bond_kr0):
    """Heavy atoms with their (at most three) bonded hydrogens and the bonds' equilibrium lengths: the constraint set
    of mc_set_hbond_constraints.  Returns (clusters (m, 4) with -1 padding, lengths (m, 3))."""
    r0 = {(int(i), int(j)): float(k[1]) for (i, j), k in zip(bonds, bond_kr0)}
    per = {}
    for (i, j) in r0:
        hi, hj = names[i].startswith("H"), names[j].startswith("H")
        if hi == hj:
            continue
        heavy, h = (j, i) if hi else (i, j)
        per.setdefault(heavy, []).append((h, r0[(i, j)]))
    clusters, lengths = [], []
    for heavy in sorted(per):
        hs = sorted(per[heavy])
        assert len(hs) <= 3, "a heavy atom carries at most three constrained hydrogens (one thread owns the whole cluster)"
        clusters.append([heavy] + [h for h, _ in hs] + [-1] * (3 - len(hs)))
        lengths.append([d for _, d in hs] + [1.0] * (3 - len(hs)))
    return np.array(clusters, np.int32).reshape(-1, 4), np.array(lengths, np.float32).reshape(-1, 3)


def bonded_globule(n_atoms=400, seed=212):
    """A C2-like globule that also carries its bonded terms: every bond of the chain graph, every
    angle and every proper dihedral it implies; nonbonded set-up as C2 (1-2/1-3 exclusions, scaled 1-4)."""
    w = globule(n_atoms, seed=seed, name=f"bonded-globule{n_atoms}")
    _, names, bonds = _saw_globule(n_atoms, seed)
    w.update(bonded_terms_from_bonds(w["xyzq"][:, :3].astype(np.float64), bonds, seed))
    w["names"] = list(names)
    w["hbond_clusters"], w["hbond_lengths"] = hbond_clusters(names, w["bonds"], w["bond_kr0"])
    return w


def solvated_c3(seed=303, n_protein=2489, n_water=7023, L=62.23):
    """C3: 2,489-atom globule + 7,023 three-site waters in a 62.23 A cubic PBC box
    (23,558 atoms, the DHFR/JAC stand-in), r_c 12 A, skin 2 A."""
    pos_p, names_p, bonds = _saw_globule(n_protein, seed, density=0.075)
    pos_p = pos_p - pos_p.mean(0) + L / 2
    rng = np.random.default_rng(seed + 11)
    q_p = rng.normal(0.0, 0.35, n_protein)
    q_p -= q_p.mean()
    m = 22
    g = (np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) * (L / m)
    # keep lattice sites whose oxygen is >= 2.9 A from every protein atom
    keep = np.ones(len(g), bool)
    for s in range(0, len(g), 512):
        d = np.linalg.norm(g[s:s + 512, None, :] - pos_p[None], axis=2).min(1)
        keep[s:s + 512] = d >= 2.9
    sites = g[keep]
    sites = sites[rng.permutation(len(sites))[:n_water]]
    n_water = len(sites)
    pos = [pos_p]
    names = list(names_p)
    charges = list(q_p)
    masses = [_CLASSES[c][2] for c in names_p]
    base = n_protein
    for o in sites:
        pos.append(_water_geometry(rng) + o)
        names += ["OW", "HW", "HW"]
        charges += [-0.834, 0.417, 0.417]
        masses += [15.999, 1.008, 1.008]
        bonds += [(base, base + 1), (base, base + 2), (base + 1, base + 2)]
        base += 3
    pos = np.concatenate(pos)
    n = len(pos)
    es, ei, p14 = topology_from_bonds(n, bonds)
    w = _pack(pos, charges, masses, names, ["H", "C", "N", "O", "S", "OW", "HW"],
              box_lo=np.zeros(3, np.float32), box_ext=np.full(3, L, np.float32), periodic=True,
              rc_lj=12.0, rc_q=12.0, skin=2.0, coul_mode=1, excl_start=es, excl_idx=ei,
              pairs14=p14, dt=0.002, name=f"C3-solvated{n}")
    maxwell_boltzmann(w["vel"], 300.0, seed + 1)
    w["bond_graph"] = [(int(i), int(j)) for i, j in bonds]
    w["mol_id"] = np.concatenate([np.zeros(n_protein, np.uint16), (1 + np.arange(n_water).repeat(3)).astype(np.uint16)])
    return w


def solvated_bonded(small=False, seed=303):
    """The C3 box with its bonded terms (flexible waters: O-H, O-H, H-H bonds and the angles they imply; the globule's bonds,
    angles and proper dihedrals) -- the periodic, decomposable counterpart of bonded_globule.  small: 600 + 3 x 1,500 atoms in
    a 40 A box with an 8 A cutoff + 1 A skin (four cell layers along z: two ranks), for the host build of the library."""
    if small:
        w = solvated_c3(seed=seed, n_protein=600, n_water=1500, L=40.0)
        w["rc_lj"] = w["rc_q"] = 8.0
        w["skin"] = 1.0
    else:
        w = solvated_c3(seed=seed)
    w["coul_mode"] = 2  # continuous at the cutoff
    w.update(bonded_terms_from_bonds(w["xyzq"][:, :3].astype(np.float64), w["bond_graph"], seed))
    w["dt"] = 0.0005
    w["name"] = "solvated-bonded%d" % len(w["xyzq"])
    return w


def lj_fluid(m=100, seed=404, seed_v=405, temp_k=86.3, name=None):
    """C4: argon LJ fluid, m^3 atoms on a jittered simple-cubic lattice at rho* = 0.8442
    (a = 3.603 A), T* = 0.72, r_c = 2.5 sigma = 8.5125 A, skin 1 A, q = 0, dt 2 fs.
    m = 100 is the 1,000,000-atom headline configuration."""
    sigma = 3.405
    a = sigma / 0.8442 ** (1 / 3)
    L = a * m
    rng = np.random.default_rng(seed)
    ax = (np.arange(m, dtype=np.float64) + 0.5) * a
    pos = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    pos += rng.uniform(-0.05 * a, 0.05 * a, pos.shape)
    n = len(pos)
    es, ei, p14 = _empty_topology(n)
    w = _pack(pos, np.zeros(n), np.full(n, 39.948), ["AR"] * n, ["AR"],
              box_lo=np.zeros(3, np.float32), box_ext=np.full(3, L, np.float32), periodic=True,
              rc_lj=2.5 * sigma, rc_q=2.5 * sigma, skin=1.0, coul_mode=0, excl_start=es, excl_idx=ei,
              pairs14=p14, dt=0.002, name=name or f"C4-ljfluid{n}")
    maxwell_boltzmann(w["vel"], temp_k, seed_v)
    return w


def ionic_mixture(m=12, seed=414, coul_mode=1):
    """A charged three-species LJ mixture at the density and cutoff of the C4 fluid (cells of ~19 atoms, rows of ~80
    entries): the dilute, short-cutoff regime of the TMA-staged force kernel with several LJ types and Coulomb terms --
    what a coarse-grained or molten-salt system looks like to the engine.  Test workload (no BASELINE config)."""
    w = lj_fluid(m=m, seed=seed, seed_v=seed + 1, name=f"ionic-mixture{m ** 3}")
    n = len(w["xyzq"])
    rng = np.random.default_rng(seed + 2)
    typ = rng.integers(0, 3, n).astype(np.uint16)
    sig = np.array([3.405, 3.0, 3.8], np.float64)
    eps = np.array([0.2381, 0.15, 0.30], np.float64)
    tab = np.zeros((3, 3, 2), np.float32)
    for a in range(3):
        for b in range(3):
            tab[a, b] = (0.5 * (sig[a] + sig[b]), np.sqrt(eps[a] * eps[b]))  # Lorentz-Berthelot
    q = np.where(typ == 1, 0.4, np.where(typ == 2, -0.4, 0.0))
    q -= q.mean()
    w["type"] = typ
    w["ljtab"] = tab
    w["xyzq"] = w["xyzq"].copy()
    w["xyzq"][:, 3] = (q * COULOMB_SCALE).astype(np.float32)
    w["coul_mode"] = coul_mode
    return w


def docking_c5(n_rec=5000, n_lig=40, n_poses=10000, seeds=(505, 506, 507)):
    """C5: receptor globule (R atoms), ligand (L atoms), rigid poses = anchors on an 8^3 grid
    within +-8 A of the site centre (site_radius default 8, reference src/docking/mod.rs:43;
    num_posits 8, docking/legacy/mod.rs:705) x seeded unit quaternions, truncated to n_poses."""
    pos_r, names_r, _ = _saw_globule(n_rec, seeds[0])
    pos_l, names_l, _ = _saw_globule(n_lig, seeds[1], density=0.04)
    rng = np.random.default_rng(seeds[2])
    q_r = rng.normal(0.0, 0.35, n_rec); q_r -= q_r.mean()
    q_l = rng.normal(0.0, 0.25, n_lig); q_l -= q_l.mean()
    order = ["H", "C", "N", "O", "S"]
    lut = {c: t for t, c in enumerate(order)}
    rec = np.zeros((n_rec, 4), np.float32); rec[:, :3] = pos_r; rec[:, 3] = q_r * COULOMB_SCALE
    lig = np.zeros((n_lig, 4), np.float32); lig[:, :3] = pos_l; lig[:, 3] = q_l * COULOMB_SCALE
    # site centre on the receptor surface
    cen = pos_r.mean(0)
    far = pos_r[np.argmax(np.linalg.norm(pos_r - cen, axis=1))]
    site = far + 3.0 * (far - cen) / np.linalg.norm(far - cen)
    site_radius, nx = 8.0, 8
    ax = -site_radius + (np.arange(nx) + 0.5) * (2 * site_radius / nx)
    anchors = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3) + site
    n_or = -(-n_poses // len(anchors))
    quats = rng.normal(size=(n_or, 4))
    quats /= np.linalg.norm(quats, axis=1, keepdims=True)
    poses = np.zeros((len(anchors) * n_or, 7), np.float32)
    poses[:, :3] = np.repeat(anchors, n_or, axis=0)
    poses[:, 3:] = np.tile(quats, (len(anchors), 1))
    poses = poses[:n_poses]
    return dict(rec=rec, rec_type=np.array([lut[c] for c in names_r], np.uint16),
                rec_hphob=np.array([c == "C" for c in names_r], np.uint8),
                lig=lig, lig_type=np.array([lut[c] for c in names_l], np.uint16),
                lig_hphob=np.array([c == "C" for c in names_l], np.uint8),
                lig_anchor=lig[0, :3].copy(), ljtab=lj_table(order), poses=poses,
                name=f"C5-dock{n_rec}x{n_lig}x{n_poses}")


def subset_box(w, n_side):
    """Helper for tests: a smaller LJ-fluid of n_side^3 atoms with the same per-atom physics."""
    return lj_fluid(m=n_side, name=f"C4-ljfluid{n_side**3}")
