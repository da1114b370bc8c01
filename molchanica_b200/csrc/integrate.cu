// integrate.cu -- velocity-Verlet kick / drift (the integrator the reference
// selects with Integrator::VerletVelocity, ui/panels/md.rs:303-305):
//     v += F/m * dt/2 * 418.4 ;  x += v * dt ;  [forces] ;  v += F/m * dt/2 * 418.4
// One fused elementwise kernel per half step; HBM-bound: 80 B/atom (x, v, F read; x, v written)
// + 16 B/atom for the reference positions of the displacement check that triggers a list rebuild
// (max displacement > skin/2).  Static atoms (AtomDynamics.static_, reference
// src/md/mod.rs:843-852) exert forces but never move.
#include "common.cuh"
#include "integrate.cuh"

namespace {

__global__ void __launch_bounds__(256) kick_drift_kernel(int n_rows, float4 *__restrict__ xyzq, float4 *__restrict__ vel,
                                                          const float4 *__restrict__ force,
                                                          const float *__restrict__ ext_force, const int *__restrict__ orig,
                                                          const uint8_t *__restrict__ flags, const float4 *__restrict__ xref,
                                                          float kick, float drift, float max_disp, float lookahead,
                                                          int *__restrict__ rebuild_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool moved_far = false;
    if (i < n_rows && !(flags[i] & MC_FLAG_STATIC)) {
        float4 v = vel[i];
        const float4 f = force[i];
        float fx = f.x, fy = f.y, fz = f.z;
        if (ext_force) {
            const int o = orig[i];
            fx += ext_force[3 * o]; fy += ext_force[3 * o + 1]; fz += ext_force[3 * o + 2];
        }
        const float s = v.w * kick * MC_ACCEL_CONV;
        v.x = fmaf(fx, s, v.x); v.y = fmaf(fy, s, v.y); v.z = fmaf(fz, s, v.z);
        vel[i] = v;
        if (drift != 0.f) {
            float4 x = xyzq[i];
            x.x = fmaf(v.x, drift, x.x); x.y = fmaf(v.y, drift, x.y); x.z = fmaf(v.z, drift, x.z);
            xyzq[i] = x;
            const float4 r = xref[i];
            const float dx = x.x - r.x, dy = x.y - r.y, dz = x.z - r.z;
            // displacement criterion (> skin/2 since the last build).  lookahead > 0 raises the flag
            // early enough that the host may act on it one step late: an upper bound (L1 norm) of the
            // distance this atom can cover in the next `lookahead` drifts is subtracted from the limit.
            const float thr = max_disp - lookahead * (fabsf(v.x) + fabsf(v.y) + fabsf(v.z)) * fabsf(drift);
            moved_far = thr <= 0.f || dx * dx + dy * dy + dz * dz > thr * thr;
            // a simulation that blew up (overlapping atoms, time step too long) must surface as an error,
            // not as out-of-range cell indices: bit 1 of the flag word reports non-finite coordinates
            if (!(fabsf(x.x) + fabsf(x.y) + fabsf(x.z) < 1.0e30f)) atomicOr(rebuild_flag, 2);
        }
    }
    if (drift != 0.f && __any_sync(MC_FULL_MASK, moved_far) && (threadIdx.x & 31) == 0) atomicOr(rebuild_flag, 1);
}

__global__ void gather_to_orig_kernel(int n, const float4 *__restrict__ sorted, const int *__restrict__ orig,
                                      float4 *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[orig[k]] = sorted[k];
}

__global__ void scatter_from_orig_kernel(int n, const float4 *__restrict__ in_orig, const int *__restrict__ orig,
                                         float4 *__restrict__ sorted, int keep_w) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float4 v = in_orig[orig[k]];
    if (keep_w) v.w = sorted[k].w;
    sorted[k] = v;
}

__global__ void l2_flush_kernel(float4 *buf, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace

void launch_kick_drift(int n_rows, float4 *xyzq, float4 *vel, const float4 *force, const float *ext_force,
                       const int *orig, const uint8_t *flags, const float4 *xref, float kick, float drift,
                       float max_disp, float lookahead, int *rebuild_flag, cudaStream_t st, int64_t *launches) {
    if (n_rows <= 0) return;
    kick_drift_kernel<<<div_up(n_rows, 256), 256, 0, st>>>(n_rows, xyzq, vel, force, ext_force, orig, flags, xref, kick,
                                                          drift, max_disp, lookahead, rebuild_flag);
    *launches += 1;
}

void launch_gather_to_orig(int n, const float4 *sorted, const int *orig, float4 *out, cudaStream_t st, int64_t *launches) {
    if (n <= 0) return;
    gather_to_orig_kernel<<<div_up(n, 256), 256, 0, st>>>(n, sorted, orig, out);
    *launches += 1;
}

void launch_scatter_from_orig(int n, const float4 *in_orig, const int *orig, float4 *sorted, int keep_w, cudaStream_t st,
                              int64_t *launches) {
    if (n <= 0) return;
    scatter_from_orig_kernel<<<div_up(n, 256), 256, 0, st>>>(n, in_orig, orig, sorted, keep_w);
    *launches += 1;
}

void launch_l2_flush(float4 *buf, size_t n_float4, cudaStream_t st, int64_t *launches) {
    l2_flush_kernel<<<1184, 256, 0, st>>>(buf, n_float4);
    *launches += 1;
}
