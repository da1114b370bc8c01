// pose_terms.h -- the rigid pose transform of the docking scan, p = anchor + R(q) (x - x_anchor), in f64 with every
// operation rounded separately (no fma contraction) and ONE rounding to f32 at the end, as the reference does
// (Pose{anchor_posit, orientation} are f64, ligand positions Vec3F32; src/docking/legacy/mod.rs:149-158, :210-214).
// The same operation sequence as dock.cu's scoring kernel, so the clash filter (dock_filter.cu, and its host twin
// mc_dock_filter_poses in dock_poses.cu) sees bit-identical points.  Shared by device and host.
#pragma once
#include <math.h>

#ifdef __CUDA_ARCH__
#define MC_POSE_HD __host__ __device__ __forceinline__
#define MC_DMUL(a, b) __dmul_rn(a, b)
#define MC_DADD(a, b) __dadd_rn(a, b)
#define MC_DSUB(a, b) __dsub_rn(a, b)
#define MC_DDIV(a, b) __ddiv_rn(a, b)
#else
#ifdef __CUDACC__
#define MC_POSE_HD __host__ __device__ __forceinline__
#else
#define MC_POSE_HD inline
#endif
// host: plain operators; the host translation units are built without fp contraction (x86-64 baseline has no fma)
#define MC_DMUL(a, b) ((a) * (b))
#define MC_DADD(a, b) ((a) + (b))
#define MC_DSUB(a, b) ((a) - (b))
#define MC_DDIV(a, b) ((a) / (b))
#endif

struct PoseQuat { double w, x, y, z; };

MC_POSE_HD PoseQuat mc_pose_quat(const float *ps /* {ax, ay, az, qw, qx, qy, qz} */) {
    double qw = ps[3], qx = ps[4], qy = ps[5], qz = ps[6];
    const double qn = sqrt(MC_DADD(MC_DADD(MC_DADD(MC_DMUL(qw, qw), MC_DMUL(qx, qx)), MC_DMUL(qy, qy)), MC_DMUL(qz, qz)));
    return PoseQuat{MC_DDIV(qw, qn), MC_DDIV(qx, qn), MC_DDIV(qy, qn), MC_DDIV(qz, qn)};
}

// l: ligand atom in its reference conformation, anchor0: the ligand's anchor point, ps: the pose.  out: posed point (f32).
MC_POSE_HD void mc_pose_point(const PoseQuat &q, const float *ps, float lx, float ly, float lz, float ax, float ay, float az, float out[3]) {
    const double vx = MC_DSUB((double)lx, (double)ax), vy = MC_DSUB((double)ly, (double)ay), vz = MC_DSUB((double)lz, (double)az);
    // v' = v + 2 (w (u x v) + u x (u x v))
    const double cx = MC_DSUB(MC_DMUL(q.y, vz), MC_DMUL(q.z, vy));
    const double cy = MC_DSUB(MC_DMUL(q.z, vx), MC_DMUL(q.x, vz));
    const double cz = MC_DSUB(MC_DMUL(q.x, vy), MC_DMUL(q.y, vx));
    const double dx = MC_DSUB(MC_DMUL(q.y, cz), MC_DMUL(q.z, cy));
    const double dy = MC_DSUB(MC_DMUL(q.z, cx), MC_DMUL(q.x, cz));
    const double dz = MC_DSUB(MC_DMUL(q.x, cy), MC_DMUL(q.y, cx));
    const double ox = MC_DADD(vx, MC_DMUL(2.0, MC_DADD(MC_DMUL(q.w, cx), dx)));
    const double oy = MC_DADD(vy, MC_DMUL(2.0, MC_DADD(MC_DMUL(q.w, cy), dy)));
    const double oz = MC_DADD(vz, MC_DMUL(2.0, MC_DADD(MC_DMUL(q.w, cz), dz)));
    out[0] = (float)MC_DADD(ox, (double)ps[0]);
    out[1] = (float)MC_DADD(oy, (double)ps[1]);
    out[2] = (float)MC_DADD(oz, (double)ps[2]);
}
