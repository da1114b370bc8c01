"""SURVEY 8a row a8 without a GPU: the pose set (make_posits_orientations / init_poses) and the clash pre-filter of
process_poses through the C ABI against the numpy restatement (oracle/dock_poses.py), plus known answers that
do not depend on either: pose counts of find_optimal_pose, anchors at cell centres, unit quaternions that map +z
onto the sampled directions, rolls that keep that direction."""
import ctypes as C

import numpy as np

from molchanica_b200 import workloads as W
from oracle import dock_poses as DP


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _make(lib, center, radius, n_pos, n_or):
    c = np.asarray(center, np.float64)
    n = C.c_int64(0)
    assert lib.mc_dock_make_poses(_ptr(c), radius, n_pos, n_or, None, 0, C.byref(n)) == 0
    out = np.zeros((n.value, 7), np.float32)
    assert lib.mc_dock_make_poses(_ptr(c), radius, n_pos, n_or, _ptr(out), n.value, C.byref(n)) == 0
    return out


def test_pose_set_matches_restatement_and_reference_counts(engine_lib):
    # find_optimal_pose: num_posits 8, num_orientations 60 -> n_lats 3, 6 longitudes, 6 rolls (legacy/mod.rs:705-706)
    assert engine_lib.mc_dock_orientation_count(60) == 108
    poses = _make(engine_lib, (10.0, -4.0, 2.5), 8.0, 8, 60)
    assert poses.shape == (512 * 108, 7)
    ref = DP.make_poses((10.0, -4.0, 2.5), 8.0, 8, 60)
    assert np.array_equal(poses[:, :3], ref[:, :3])
    assert np.abs(poses[:, 3:] - ref[:, 3:]).max() < 2e-6     # libm vs numpy f32 acos / sin / cos
    # anchors: cell centres of the 8^3 grid, x slowest, one block of 108 poses per anchor
    a = poses[::108, :3]
    assert np.allclose(a[0], [10 - 8 + 1.0, -4 - 8 + 1.0, 2.5 - 8 + 1.0]) and np.allclose(a[1] - a[0], [0, 0, 2.0])
    assert np.allclose(a[64] - a[0], [2.0, 0, 0]) and np.allclose(a.mean(0), [10.0, -4.0, 2.5], atol=1e-5)
    # too small a buffer is an error, not a partial write
    n = C.c_int64(0)
    small = np.zeros((10, 7), np.float32)
    assert engine_lib.mc_dock_make_poses(_ptr(np.zeros(3)), 8.0, 8, 60, _ptr(small), 10, C.byref(n)) == -4


def test_orientations_are_rolls_about_equal_area_directions(engine_lib):
    q = _make(engine_lib, (0, 0, 0), 1.0, 1, 60)[:, 3:].astype(np.float64)
    assert np.allclose((q * q).sum(1), 1.0, atol=1e-6)
    w, v = q[:, 0], q[:, 1:]
    z = np.array([0.0, 0.0, 1.0])
    rz = z + 2.0 * (w[:, None] * np.cross(v, z) + np.cross(v, np.cross(v, z)))   # where each pose sends +z
    # 6 rolls share a direction; the 18 directions sit on three latitude bands mu = -2/3, 0, 2/3, six longitudes each
    d = rz.reshape(18, 6, 3)
    assert np.abs(d - d[:, :1]).max() < 1e-6
    assert np.allclose(sorted(set(np.round(d[:, 0, 2], 5))), [-2 / 3, 0.0, 2 / 3], atol=1e-5)
    lon = np.degrees(np.arctan2(d[:6, 0, 1], d[:6, 0, 0])) % 360
    assert np.allclose(np.sort(lon), [30, 90, 150, 210, 270, 330], atol=1e-3)
    # the rolls of one direction differ by 60 degrees about it
    rel = q.reshape(18, 6, 4)
    x = np.array([1.0, 0.0, 0.0])
    for k in range(18):
        wk, vk = rel[k, :, 0], rel[k, :, 1:]
        rx = x + 2.0 * (wk[:, None] * np.cross(vk, x) + np.cross(vk, np.cross(vk, x)))
        cosang = np.clip((rx[1:] * rx[:-1]).sum(1) - (rx[1:] @ d[k, 0]) * (rx[:-1] @ d[k, 0]), -1, 1)
        perp = 1.0 - (rx[0] @ d[k, 0]) ** 2
        assert np.allclose(cosang / perp, np.cos(np.pi / 3), atol=1e-5)


def test_near_site_and_clash_filter_match_restatement(engine_lib):
    d = W.docking_c5(n_rec=1500, n_lig=24, n_poses=64, seeds=(515, 516, 517))
    rec, lig = d["rec"], d["lig"]
    site = rec[:, :3].astype(np.float64).mean(0) + np.array([6.0, 0.0, 0.0])
    hetero = (np.arange(len(rec)) % 17 == 0).astype(np.uint8)
    n = C.c_int64(0)
    idx = np.zeros(len(rec), np.int32)
    assert engine_lib.mc_dock_near_site(len(rec), _ptr(rec), _ptr(hetero), _ptr(site), 8.0, _ptr(idx), C.byref(n)) == 0
    want = DP.near_site(rec, hetero, site, 8.0)
    assert n.value == len(want) > 50 and np.array_equal(idx[:n.value], want)
    near = np.ascontiguousarray(rec[want])
    near_c = np.ascontiguousarray(d["rec_hphob"][want])      # carbon flags of the workload
    poses = _make(engine_lib, site, 8.0, 4, 60)
    keep = np.zeros(len(poses), np.uint8)
    kept = C.c_int64(0)
    anchor = np.ascontiguousarray(d["lig_anchor"], np.float32)
    assert engine_lib.mc_dock_filter_poses(len(near), _ptr(near), _ptr(near_c), len(lig), _ptr(lig), _ptr(d["lig_hphob"]),
                                           _ptr(anchor), 1.7, len(poses), _ptr(poses), _ptr(keep), C.byref(kept)) == 0
    ref = DP.filter_poses(near, near_c, lig, d["lig_hphob"], anchor, poses, 1.7)
    assert np.array_equal(keep, ref)
    assert 0 < kept.value == int(ref.sum()) < len(poses)      # the filter does remove some and keep some


def test_flexible_pose_set_and_downstream_masks(engine_lib):
    """init_poses with flexible bonds (legacy/mod.rs:453-500) and the two sides of a rotatable bond, against the numpy
    restatement and known answers: pose count = rigid x angles^bonds, first bond slowest, linspace end points 0 and TAU;
    a chain splits at the bond, a ring bond is refused."""
    c = np.asarray((1.0, 2.0, 3.0), np.float64)
    n = C.c_int64(0)
    assert engine_lib.mc_dock_make_poses_flex(_ptr(c), 6.0, 2, 16, 2, 3, None, 0, C.byref(n)) == 0
    n_rigid = 8 * engine_lib.mc_dock_orientation_count(16)
    assert n.value == n_rigid * 9
    out = np.zeros((n.value, 9), np.float32)
    assert engine_lib.mc_dock_make_poses_flex(_ptr(c), 6.0, 2, 16, 2, 3, _ptr(out), n.value, C.byref(n)) == 0
    ref = DP.make_poses_flex(c, 6.0, 2, 16, 2, 3)
    assert np.array_equal(out, ref)
    assert np.array_equal(out[:9, 7], np.repeat(np.float32([0.0, np.pi, 2 * np.pi]), 3))    # first bond slowest
    assert np.array_equal(out[:9, 8], np.tile(np.float32([0.0, np.pi, 2 * np.pi]), 3))
    assert np.array_equal(out[:9, :7], np.repeat(out[:1, :7], 9, axis=0))
    # rigid case = mc_dock_make_poses
    assert engine_lib.mc_dock_make_poses_flex(_ptr(c), 6.0, 2, 16, 0, 0, None, 0, C.byref(n)) == 0 and n.value == n_rigid
    # a 6-chain with a 3-ring at its end: 0-1-2-3-4-5, 3-5
    bonds = np.array([[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [3, 5]], np.int32)
    flex = np.array([1, 2], np.int32)
    axis, mask = np.zeros((2, 2), np.int32), np.zeros((2, 6), np.uint8)
    assert engine_lib.mc_dock_flex_masks(6, len(bonds), _ptr(bonds), 2, _ptr(flex), _ptr(axis), _ptr(mask)) == 0
    ra, rm = DP.flex_masks(6, bonds, flex)
    assert np.array_equal(axis, ra) and np.array_equal(mask, rm)
    assert mask[0].tolist() == [0, 0, 0, 1, 1, 1] and mask[1].tolist() == [0, 0, 0, 0, 1, 1] and axis.tolist() == [[1, 2], [2, 3]]
    ring = np.array([3], np.int32)   # bond 3-4 lies in the ring 3-4-5
    assert engine_lib.mc_dock_flex_masks(6, len(bonds), _ptr(bonds), 1, _ptr(ring), _ptr(axis), _ptr(mask)) != 0


def test_torsion_restatement_known_answers():
    """apply_torsions: a half turn about the z axis maps (x, y, z) to (-x, -y, z) for downstream atoms only; bond lengths to
    the axis atoms are kept; a second torsion acts on the already rotated atoms."""
    lig = np.array([[0, 0, 0], [0, 0, 1.5], [1.0, 0.5, 2.0], [1.0, 0.5, 3.5], [2.2, 0.5, 3.9]], np.float64)
    axis, mask = DP.flex_masks(5, [[0, 1], [1, 2], [2, 3], [3, 4]], [0])
    x = DP.apply_torsions(lig, axis, mask, [np.pi])
    assert np.allclose(x[:2], lig[:2]) and np.allclose(x[2:, 0], -lig[2:, 0]) and np.allclose(x[2:, 1], -lig[2:, 1]) and np.allclose(x[2:, 2], lig[2:, 2])
    axis2, mask2 = DP.flex_masks(5, [[0, 1], [1, 2], [2, 3], [3, 4]], [0, 2])
    y = DP.apply_torsions(lig, axis2, mask2, [0.7, -1.1])
    d = lambda p, i, j: np.linalg.norm(p[i] - p[j])
    for i, j in ((0, 1), (1, 2), (2, 3), (3, 4)):
        assert abs(d(y, i, j) - d(lig, i, j)) < 1e-12
    assert not np.allclose(y[4], DP.apply_torsions(lig, axis2[:1], mask2[:1], [0.7])[4])
