// dock_poses.cu -- SURVEY 8a row a8: the pose set of the docking scan and its cheap geometric pre-filter,
// host side (the reference runs both serially on the CPU before the rayon scoring loop).  No device code:
// these entry points work without a GPU and feed mc_dock_score.
//
//   mc_dock_make_poses   make_posits_orientations + init_poses for a rigid ligand
//                        (reference src/docking/legacy/mod.rs:386-450, :453-500)
//   mc_dock_near_site    find_rec_atoms_near_site (legacy/prep.rs:506-532, ATOM_NEAR_SITE_DIST_THRESH = 1.4, mod.rs:68)
//   mc_dock_filter_poses the clash pre-filter of process_poses (legacy/mod.rs:522-573): a pose is dropped when a
//                        sampled ligand carbon (every 4th ligand atom, prep.rs:22) comes closer than 1.1 x the van
//                        der Waals radius to a sampled receptor carbon (every 6th near-site atom, prep.rs:21,133)
//
// Quaternion helpers restate lin_alg 1.4.3 (Cargo.toml:19; crate not vendored) [EXTERNAL-RECALL]:
// from_unit_vecs(a, b) = normalise(1 + a.b, a x b), from_axis_angle(axis, t) = (cos t/2, axis sin t/2),
// Hamilton product.  The pose transform is the one dock.cu applies (p = anchor + R(q)(x - x_anchor), f64, rounded
// once to f32), so the filter sees bit-identical points to the scoring kernel's.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/molchanica_md.h"
#include "pose_terms.h"

namespace {

struct Quat { double w, x, y, z; };

Quat q_mul(const Quat &a, const Quat &b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}

Quat q_normalized(const Quat &q) {
    const double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    return {q.w / n, q.x / n, q.y / n, q.z / n};
}

Quat q_from_unit_vecs(const double a[3], const double b[3]) {
    const double d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    if (d < -1.0 + 1e-12) return {0.0, 1.0, 0.0, 0.0};  // antiparallel: half turn about x (a = +z here)
    const Quat q = {1.0 + d, a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    return q_normalized(q);
}

Quat q_from_axis_angle(const double axis[3], double angle) {
    const double s = std::sin(angle * 0.5);
    return {std::cos(angle * 0.5), axis[0] * s, axis[1] * s, axis[2] * s};
}

const float TAU_F = 6.28318530717958647692f;

}  // namespace

extern "C" int mc_dock_orientation_count(int num_orientations) {
    // n_lats = floor((num_orientations / 2)^(1/3)) in f32, n_lons = n_rolls = 2 n_lats  (legacy/mod.rs:421-425)
    const int n_lats = (int)std::pow((float)num_orientations / 2.f, 1.f / 3.f);
    return n_lats * (2 * n_lats) * (2 * n_lats);
}

extern "C" int mc_dock_make_poses(const double site_center[3], double site_radius, int num_posits, int num_orientations,
                                  float *out_poses, int64_t cap, int64_t *n_out) {
    if (!site_center || !n_out || num_posits < 1 || num_orientations < 2 || !(site_radius > 0.0)) return MC_E_INVALID;
    const int n = num_posits;
    const int n_lats = (int)std::pow((float)num_orientations / 2.f, 1.f / 3.f);
    const int n_lons = n_lats * 2, n_rolls = n_lons;
    const int64_t n_anchor = (int64_t)n * n * n, n_or = (int64_t)n_lats * n_lons * n_rolls;
    *n_out = n_anchor * n_or;
    if (!out_poses) return MC_OK;
    if (cap < *n_out) return MC_E_CAPACITY;
    // orientations: latitude bands equal in mu = cos(phi), longitudes, then rolls about the direction (f32 angles,
    // f64 quaternions, as the reference mixes them)
    std::vector<Quat> ors;
    ors.reserve((size_t)n_or);
    const double zaxis[3] = {0.0, 0.0, 1.0};
    for (int i_lat = 0; i_lat < n_lats; ++i_lat) {
        const float frac = ((float)i_lat + 0.5f) / (float)n_lats;
        const float mu = -1.0f + 2.0f * frac;
        const float phi = std::acos(mu);
        for (int i_lon = 0; i_lon < n_lons; ++i_lon) {
            const float theta = ((float)i_lon + 0.5f) * TAU_F / (float)n_lons;
            float v[3] = {std::sin(phi) * std::cos(theta), std::sin(phi) * std::sin(theta), mu};
            const float len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            const double dir[3] = {(double)(v[0] / len), (double)(v[1] / len), (double)(v[2] / len)};
            const Quat orq = q_from_unit_vecs(zaxis, dir);
            for (int roll = 0; roll < n_rolls; ++roll) {
                const float angle = (float)roll * TAU_F / (float)n_rolls;
                ors.push_back(q_mul(q_from_axis_angle(dir, (double)angle), orq));
            }
        }
    }
    // anchors: cell centres of an n^3 grid over the cube of half-width site_radius, x slowest (i, j, k loops)
    const double d = 2.0 * site_radius / (double)n;
    int64_t p = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            for (int k = 0; k < n; ++k) {
                const double ax = site_center[0] - site_radius + ((double)i + 0.5) * d;
                const double ay = site_center[1] - site_radius + ((double)j + 0.5) * d;
                const double az = site_center[2] - site_radius + ((double)k + 0.5) * d;
                for (const Quat &q : ors) {
                    float *o = out_poses + 7 * p++;
                    o[0] = (float)ax; o[1] = (float)ay; o[2] = (float)az;
                    o[3] = (float)q.w; o[4] = (float)q.x; o[5] = (float)q.y; o[6] = (float)q.z;
                }
            }
    return MC_OK;
}

// init_poses with flexible bonds (legacy/mod.rs:453-500): every rigid pose (anchor x orientation) times the cartesian
// product of angles_per_bond dihedral angles per flexible bond, angles = linspace(0, TAU, angles_per_bond) (end points
// included, as the reference's linspace gives them); the FIRST bond varies slowest, the last fastest (the reference
// extends every existing combination by all angles of the next bond).  Row: {ax, ay, az, qw, qx, qy, qz, t_0 .. t_{F-1}}.
extern "C" int mc_dock_make_poses_flex(const double site_center[3], double site_radius, int num_posits, int num_orientations,
                                       int n_flex_bonds, int angles_per_bond, float *out_poses, int64_t cap, int64_t *n_out) {
    if (!n_out || n_flex_bonds < 0 || n_flex_bonds > MC_DOCK_MAX_FLEX || (n_flex_bonds > 0 && angles_per_bond < 1)) return MC_E_INVALID;
    int64_t n_rigid = 0;
    int rc = mc_dock_make_poses(site_center, site_radius, num_posits, num_orientations, nullptr, 0, &n_rigid);
    if (rc != MC_OK) return rc;
    int64_t combos = 1;
    for (int b = 0; b < n_flex_bonds; ++b) {
        combos *= angles_per_bond;
        if (combos > ((int64_t)1 << 40) / std::max<int64_t>(n_rigid, 1)) return MC_E_CAPACITY;
    }
    *n_out = n_rigid * combos;
    if (!out_poses) return MC_OK;
    if (cap < *n_out) return MC_E_CAPACITY;
    std::vector<float> rigid((size_t)7 * n_rigid);
    if ((rc = mc_dock_make_poses(site_center, site_radius, num_posits, num_orientations, rigid.data(), n_rigid, &n_rigid)) != MC_OK) return rc;
    std::vector<float> angles((size_t)std::max(angles_per_bond, 1));
    for (int a = 0; a < angles_per_bond; ++a)  // linspace(0., TAU, n): f32, both ends
        angles[(size_t)a] = angles_per_bond == 1 ? 0.f : (float)a * (TAU_F / (float)(angles_per_bond - 1));
    const int stride = 7 + n_flex_bonds;
    for (int64_t r = 0; r < n_rigid; ++r)
        for (int64_t k = 0; k < combos; ++k) {
            float *o = out_poses + (size_t)stride * (size_t)(r * combos + k);
            memcpy(o, rigid.data() + 7 * r, 7 * sizeof(float));
            int64_t rem = k;
            for (int b = n_flex_bonds - 1; b >= 0; --b) {  // last bond fastest
                o[7 + b] = angles[(size_t)(rem % angles_per_bond)];
                rem /= angles_per_bond;
            }
        }
    return MC_OK;
}

// The two sides of every flexible bond (the reference's rotate_around_bond, EXTERNAL crate: "divide all atoms into those
// upstream of this bond and those downstream; rotate all downstream atoms", mol_alignment.rs:114-122): the bond graph is
// cut at bond (a0, a1), everything still connected to a1 is downstream.  axis_out: {a0, a1} per flexible bond;
// mask_out[f * n_lig + a] = 1 for downstream atoms (a1 itself lies on the axis and is left out).  A bond inside a ring
// (a0 reachable from a1 after the cut) cannot rotate: MC_E_INVALID.
extern "C" int mc_dock_flex_masks(int64_t n_lig, int64_t n_bonds, const int32_t *bonds, int n_flex_bonds, const int32_t *flex_bond_idx,
                                  int32_t *axis_out, uint8_t *mask_out) {
    if (n_lig <= 0 || n_bonds < 0 || (n_bonds > 0 && !bonds) || n_flex_bonds < 0 || n_flex_bonds > MC_DOCK_MAX_FLEX ||
        (n_flex_bonds > 0 && (!flex_bond_idx || !axis_out || !mask_out)))
        return MC_E_INVALID;
    std::vector<std::vector<int32_t>> adj((size_t)n_lig);
    for (int64_t b = 0; b < n_bonds; ++b) {
        const int32_t u = bonds[2 * b], v = bonds[2 * b + 1];
        if (u < 0 || v < 0 || u >= n_lig || v >= n_lig || u == v) return MC_E_INVALID;
        adj[(size_t)u].push_back(v);
        adj[(size_t)v].push_back(u);
    }
    for (int f = 0; f < n_flex_bonds; ++f) {
        const int32_t bi = flex_bond_idx[f];
        if (bi < 0 || bi >= n_bonds) return MC_E_INVALID;
        const int32_t a0 = bonds[2 * bi], a1 = bonds[2 * bi + 1];
        axis_out[2 * f] = a0;
        axis_out[2 * f + 1] = a1;
        uint8_t *m = mask_out + (size_t)f * (size_t)n_lig;
        memset(m, 0, (size_t)n_lig);
        std::vector<int32_t> stack{a1};
        std::vector<uint8_t> seen((size_t)n_lig, 0);
        seen[(size_t)a1] = 1;
        while (!stack.empty()) {
            const int32_t u = stack.back();
            stack.pop_back();
            for (int32_t v : adj[(size_t)u]) {
                if ((u == a1 && v == a0) || seen[(size_t)v]) continue;  // the cut bond
                if (v == a0) return MC_E_INVALID;                       // ring: a0 reached the other way round
                seen[(size_t)v] = 1;
                m[(size_t)v] = 1;
                stack.push_back(v);
            }
        }
    }
    return MC_OK;
}

extern "C" int mc_dock_near_site(int64_t n_rec, const mc_float4 *rec_xyzq, const uint8_t *rec_hetero, const double site_center[3],
                                 double site_radius, int32_t *out_idx, int64_t *n_out) {
    if (!rec_xyzq || !site_center || !n_out || n_rec < 0) return MC_E_INVALID;
    const double thresh = 1.4 * site_radius;  // ATOM_NEAR_SITE_DIST_THRESH (legacy/mod.rs:68)
    int64_t m = 0;
    for (int64_t i = 0; i < n_rec; ++i) {
        const double dx = (double)rec_xyzq[i].x - site_center[0], dy = (double)rec_xyzq[i].y - site_center[1],
                     dz = (double)rec_xyzq[i].z - site_center[2];
        if (std::sqrt(dx * dx + dy * dy + dz * dz) < thresh && !(rec_hetero && rec_hetero[i])) {
            if (out_idx) out_idx[m] = (int32_t)i;
            ++m;
        }
    }
    *n_out = m;
    return MC_OK;
}

extern "C" int mc_dock_filter_poses(int64_t n_rec, const mc_float4 *rec_xyzq, const uint8_t *rec_is_carbon, int64_t n_lig,
                                    const mc_float4 *lig_xyzq, const uint8_t *lig_is_carbon, const float lig_anchor[3],
                                    float vdw_radius, int64_t n_poses, const float *poses, uint8_t *keep, int64_t *n_kept) {
    if (!rec_xyzq || !rec_is_carbon || !lig_xyzq || !lig_is_carbon || !lig_anchor || n_rec < 0 || n_lig < 0 || n_poses < 0 ||
        (n_poses > 0 && (!poses || !keep)))
        return MC_E_INVALID;
    std::vector<int64_t> rs, ls;
    for (int64_t i = 0; i < n_rec; ++i)
        if (rec_is_carbon[i] && i % 6 == 0) rs.push_back(i);  // REC_SAMPLE_RATIO (prep.rs:21,133)
    for (int64_t i = 0; i < n_lig; ++i)
        if (lig_is_carbon[i] && i % 4 == 0) ls.push_back(i);  // LIGAND_SAMPLE_RATIO (prep.rs:22, mod.rs:540)
    const float limit = vdw_radius * 1.1f;
    std::vector<float> lp(3 * ls.size());
    int64_t kept = 0;
    for (int64_t p = 0; p < n_poses; ++p) {
        const float *ps = poses + 7 * p;
        const PoseQuat q = mc_pose_quat(ps);
        for (size_t a = 0; a < ls.size(); ++a) {
            const mc_float4 &l = lig_xyzq[ls[a]];
            mc_pose_point(q, ps, l.x, l.y, l.z, lig_anchor[0], lig_anchor[1], lig_anchor[2], &lp[3 * a]);
        }
        bool clash = false;
        for (size_t r = 0; r < rs.size() && !clash; ++r) {
            const mc_float4 &ra = rec_xyzq[rs[r]];
            for (size_t a = 0; a < ls.size(); ++a) {
                const float ex = ra.x - lp[3 * a], ey = ra.y - lp[3 * a + 1], ez = ra.z - lp[3 * a + 2];
                if (std::sqrt(ex * ex + ey * ey + ez * ez) < limit) { clash = true; break; }
            }
        }
        keep[p] = clash ? 0 : 1;
        kept += clash ? 0 : 1;
    }
    if (n_kept) *n_kept = kept;
    return MC_OK;
}
