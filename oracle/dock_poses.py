"""CPU restatement (numpy) of the docking pose set and its clash pre-filter -- SURVEY 8a row a8.

TEST INFRASTRUCTURE ONLY (see oracle/md_oracle.c): imported by tests/, never by the product package.
Parity unpinned: the reference holds no test or fixture for this path (src/tests.rs is empty) and cannot be
built here (no Rust toolchain); each function follows the reference lines it cites, and the quaternion
helpers restate the un-vendored crate lin_alg 1.4.3 (Cargo.toml:19) [EXTERNAL-RECALL].

    make_posits_orientations  reference src/docking/legacy/mod.rs:386-450
    init_poses (rigid ligand) reference src/docking/legacy/mod.rs:453-500
    near_site                 reference src/docking/legacy/prep.rs:506-532 (threshold 1.4 x site_radius, mod.rs:68)
    filter_poses              reference src/docking/legacy/mod.rs:522-573 (+ sampling ratios prep.rs:21-22,133)
"""
import numpy as np

f32 = np.float32
TAU32 = f32(6.28318530717958647692)


def _q_mul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                     a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]], np.float64)


def _q_from_unit_vecs(a, b):
    d = float(np.dot(a, b))
    if d < -1.0 + 1e-12:
        return np.array([0.0, 1.0, 0.0, 0.0])
    q = np.array([1.0 + d, *np.cross(a, b)], np.float64)
    return q / np.sqrt(np.sum(q * q))


def _q_from_axis_angle(axis, angle):
    s = np.sin(angle * 0.5)
    return np.array([np.cos(angle * 0.5), axis[0] * s, axis[1] * s, axis[2] * s], np.float64)


def orientations(num_orientations):
    """mod.rs:421-447: latitude bands equal in mu, 2 n_lats longitudes, 2 n_lats rolls about the direction."""
    n_lats = int(f32(num_orientations / 2.0) ** f32(1.0 / 3.0))
    n_lons = n_rolls = 2 * n_lats
    out = []
    z = np.array([0.0, 0.0, 1.0])
    for i_lat in range(n_lats):
        frac = (f32(i_lat) + f32(0.5)) / f32(n_lats)
        mu = f32(-1.0) + f32(2.0) * frac
        phi = np.arccos(mu, dtype=f32)
        for i_lon in range(n_lons):
            theta = (f32(i_lon) + f32(0.5)) * TAU32 / f32(n_lons)
            v = np.array([np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), mu], f32)
            v = v / np.sqrt(np.sum(v * v), dtype=f32)
            d = v.astype(np.float64)
            orq = _q_from_unit_vecs(z, d)
            for roll in range(n_rolls):
                angle = f32(roll) * TAU32 / f32(n_rolls)
                out.append(_q_mul(_q_from_axis_angle(d, float(angle)), orq))
    return np.array(out, np.float64)


def make_poses(site_center, site_radius, num_posits, num_orientations):
    """n x {ax, ay, az, qw, qx, qy, qz} as f32, anchor-major (init_poses iterates anchors outside orientations)."""
    n = num_posits
    c = np.asarray(site_center, np.float64)
    d = 2.0 * site_radius / n
    ax = np.array([[c[0] - site_radius + (i + 0.5) * d, c[1] - site_radius + (j + 0.5) * d, c[2] - site_radius + (k + 0.5) * d]
                   for i in range(n) for j in range(n) for k in range(n)], np.float64)
    ors = orientations(num_orientations)
    poses = np.zeros((len(ax) * len(ors), 7), f32)
    poses[:, :3] = np.repeat(ax, len(ors), axis=0).astype(f32)
    poses[:, 3:] = np.tile(ors, (len(ax), 1)).astype(f32)
    return poses


def make_poses_flex(site_center, site_radius, num_posits, num_orientations, n_flex_bonds, angles_per_bond):
    """init_poses with flexible bonds (mod.rs:453-500): every rigid pose times the cartesian product of
    linspace(0, TAU, angles_per_bond) per bond, first bond slowest.  n x (7 + n_flex_bonds) f32."""
    import itertools
    rigid = make_poses(site_center, site_radius, num_posits, num_orientations)
    if n_flex_bonds == 0:
        return rigid
    angles = [f32(0.0)] if angles_per_bond == 1 else [f32(a) * (TAU32 / f32(angles_per_bond - 1)) for a in range(angles_per_bond)]
    combos = np.array(list(itertools.product(angles, repeat=n_flex_bonds)), f32)
    out = np.zeros((len(rigid) * len(combos), 7 + n_flex_bonds), f32)
    out[:, :7] = np.repeat(rigid, len(combos), axis=0)
    out[:, 7:] = np.tile(combos, (len(rigid), 1))
    return out


def flex_masks(n_lig, bonds, flex_bond_idx):
    """Downstream side of every flexible bond: atoms still connected to a1 when bond (a0, a1) is cut (a1 itself excluded).
    Returns (axis (F, 2), mask (F, n_lig)); raises ValueError for a bond inside a ring."""
    bonds = np.asarray(bonds, np.int64).reshape(-1, 2)
    adj = [[] for _ in range(n_lig)]
    for u, v in bonds:
        adj[u].append(v)
        adj[v].append(u)
    axis = np.zeros((len(flex_bond_idx), 2), np.int32)
    mask = np.zeros((len(flex_bond_idx), n_lig), np.uint8)
    for f, bi in enumerate(flex_bond_idx):
        a0, a1 = bonds[bi]
        axis[f] = (a0, a1)
        seen, todo = {int(a1)}, [int(a1)]
        while todo:
            u = todo.pop()
            for v in adj[u]:
                if (u == a1 and v == a0) or v in seen:
                    continue
                if v == a0:
                    raise ValueError("flexible bond inside a ring")
                seen.add(int(v))
                mask[f, v] = 1
                todo.append(int(v))
    return axis, mask


def apply_torsions(lig_xyz, axis, mask, angles):
    """The conformer of one pose in f64: for every flexible bond in order, the downstream atoms are rotated by its angle
    about the axis a0 -> a1 through a1 (Rodrigues), a later axis seeing the atoms where earlier rotations left them."""
    x = np.asarray(lig_xyz, np.float64)[:, :3].copy()
    for (a0, a1), m, t in zip(axis, mask, angles):
        p1 = x[a1].copy()
        u = p1 - x[a0]
        u /= np.sqrt((u * u).sum())
        c, s = np.cos(float(t)), np.sin(float(t))
        sel = np.nonzero(m)[0]
        v = x[sel] - p1
        x[sel] = p1 + v * c + np.cross(u, v) * s + np.outer((v @ u) * (1.0 - c), u)
    return x


def pose_points_flex(lig_xyz, lig_anchor, pose, axis, mask):
    """Torsions (f64) then the rigid transform (f64), rounded once to f32."""
    conf = apply_torsions(lig_xyz[:, :3].astype(f32), axis, mask, pose[7:])
    q = pose[3:7].astype(np.float64)
    q = q / np.sqrt(np.sum(q * q))
    v = conf - np.asarray(lig_anchor, f32).astype(np.float64)
    qv = q[1:]
    c = np.cross(qv, v)
    dd = np.cross(qv, c)
    return (v + 2.0 * (q[0] * c + dd) + pose[:3].astype(np.float64)).astype(f32)


def near_site(rec_xyz, hetero, site_center, site_radius):
    d = np.sqrt(((rec_xyz[:, :3].astype(np.float64) - np.asarray(site_center, np.float64)) ** 2).sum(1))
    ok = d < 1.4 * site_radius
    if hetero is not None:
        ok &= ~np.asarray(hetero, bool)
    return np.nonzero(ok)[0].astype(np.int32)


def pose_points(lig_xyz, lig_anchor, pose):
    """The ligand under one pose: anchor + R(q)(x - x_anchor) in f64, rounded once to f32 (mod.rs:149-158, the
    transform oracle dock_score and dock.cu share)."""
    q = pose[3:].astype(np.float64)
    q = q / np.sqrt(np.sum(q * q))
    v = lig_xyz[:, :3].astype(np.float64) - np.asarray(lig_anchor, f32).astype(np.float64)
    qv = q[1:]
    c = np.cross(qv, v)
    dd = np.cross(qv, c)
    return (v + 2.0 * (q[0] * c + dd) + pose[:3].astype(np.float64)).astype(f32)


def filter_poses(rec_xyz, rec_is_carbon, lig_xyz, lig_is_carbon, lig_anchor, poses, vdw_radius=1.7):
    rs = [i for i in range(len(rec_xyz)) if rec_is_carbon[i] and i % 6 == 0]
    ls = [i for i in range(len(lig_xyz)) if lig_is_carbon[i] and i % 4 == 0]
    limit = f32(vdw_radius) * f32(1.1)
    r = rec_xyz[rs, :3].astype(f32)
    keep = np.ones(len(poses), np.uint8)
    for p, pose in enumerate(poses):
        pts = pose_points(lig_xyz[ls], lig_anchor, pose)
        e = r[:, None, :] - pts[None, :, :]
        dist = np.sqrt((e[..., 0] * e[..., 0] + e[..., 1] * e[..., 1]) + e[..., 2] * e[..., 2], dtype=f32)
        if (dist < limit).any():
            keep[p] = 0
    return keep
