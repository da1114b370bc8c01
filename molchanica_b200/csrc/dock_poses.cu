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
