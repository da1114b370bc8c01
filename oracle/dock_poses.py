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
