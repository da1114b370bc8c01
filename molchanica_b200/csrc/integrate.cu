// integrate.cu -- velocity-Verlet kick / drift (the integrator the reference
// selects with Integrator::VerletVelocity, ui/panels/md.rs:303-305):
//     v += F/m * dt/2 * 418.4 ;  x += v * dt ;  [forces] ;  v += F/m * dt/2 * 418.4
// One fused elementwise kernel per half step; HBM-bound: 80 B/atom (x, v, F read; x, v written)
// + 16 B/atom for the reference positions of the displacement check that triggers a list rebuild
// (max displacement > skin/2).  Static atoms (AtomDynamics.static_, reference
// src/md/mod.rs:843-852) exert forces but never move.
#include "common.cuh"
#include "integrate.cuh"
#include "halo_sync.cuh"

namespace {

// HALO: the decomposed step (comm.cu).  Block 0 first tells both neighbours that every kernel of this
// rank that read the ghosts of the previous epochs has finished (they precede this launch in the
// stream); blocks that own boundary-layer rows wait for the matching ack of the target rank, then
// every new position of a boundary layer is also stored into that rank's ghost block through the
// mapped peer pointer; the last block to finish publishes the epoch in the neighbours' ready flags.
template <bool HALO>
__global__ void __launch_bounds__(256) kick_drift_kernel(int n_rows, float4 *__restrict__ xyzq, float4 *__restrict__ vel,
                                                          const float4 *__restrict__ force,
                                                          const float *__restrict__ ext_force, const int *__restrict__ orig,
                                                          const uint8_t *__restrict__ flags, const float4 *__restrict__ xref,
                                                          float kick, float drift, float max_disp, float lookahead,
                                                          int *__restrict__ rebuild_flag, const HaloPush hp,
                                                          uint32_t *__restrict__ done_counter, volatile int *host_flag,
                                                          int step_tag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool moved_far = false;
    float disp2 = 0.f;
    const bool push = HALO && hp.to_prev != nullptr;
    // interior rows first, the two boundary layers last: by the time the pushing blocks are scheduled the
    // neighbours' acks have usually arrived, and only those blocks pay for the system-scope fence below
    const int n_interior = HALO ? hp.last_begin - hp.n_first : 0;
    const bool push_block = push && (int)((blockIdx.x + 1) * blockDim.x) > n_interior;
    if (HALO) {
        if (i < n_rows) i = i < n_interior ? i + hp.n_first : (i < hp.last_begin ? i - n_interior : i);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            halo_st_release_sys(hp.sig_ack_prev, hp.epoch - 1u);
            halo_st_release_sys(hp.sig_ack_next, hp.epoch - 1u);
        }
        if (push_block) {
            if (threadIdx.x == 0) {
                halo_spin(hp.ack_prev, hp.epoch - 1u, hp.err);
                halo_spin(hp.ack_next, hp.epoch - 1u, hp.err);
            }
            __syncthreads();
        }
    }
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
            if (push) {
                if (i < hp.n_first) hp.to_prev[i] = x;
                if (i >= hp.last_begin) hp.to_next[i - hp.last_begin] = x;
            }
            const float4 r = xref[i];
            const float dx = x.x - r.x, dy = x.y - r.y, dz = x.z - r.z;
            // displacement criterion (> skin/2 since the last build).  lookahead > 0 raises the flag
            // early enough that the host may act on it one step late: the
            // distance this atom can cover in the next `lookahead` drifts is subtracted from the limit.
            const float speed = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
            const float thr = max_disp - lookahead * speed * fabsf(drift);
            disp2 = dx * dx + dy * dy + dz * dz;
            moved_far = thr <= 0.f || disp2 > thr * thr;
            // a simulation that blew up (overlapping atoms, time step too long) must surface as an error,
            // not as out-of-range cell indices: bit 1 of the flag word reports non-finite coordinates
            if (!(fabsf(x.x) + fabsf(x.y) + fabsf(x.z) < 1.0e30f)) atomicOr(rebuild_flag, 2);
        }
    }
    if (drift != 0.f && __any_sync(MC_FULL_MASK, moved_far) && (threadIdx.x & 31) == 0) atomicOr(rebuild_flag, 1);
    if (HALO && drift != 0.f) {
        // largest squared displacement since the last build: one atomic per BLOCK, and only when it raises the
        // running maximum (thousands of same-address atomics per launch would cost more than the kernel itself)
        __shared__ float s_d2[8];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) disp2 = fmaxf(disp2, __shfl_xor_sync(MC_FULL_MASK, disp2, d));
        if ((threadIdx.x & 31) == 0) s_d2[threadIdx.x >> 5] = disp2;
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = s_d2[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) m = fmaxf(m, s_d2[k]);
            if (m > __int_as_float(*reinterpret_cast<volatile int *>(hp.max_disp2)))
                atomicMax(hp.max_disp2, __float_as_int(m));  // non-negative floats order like ints
        }
    }
    if (!HALO && host_flag) {
        // The last block to finish publishes {step tag, flag bits} straight into pinned host memory: the host
        // polls that word instead of queueing a copy + event between this kernel and the pair kernel.
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(done_counter, 1u) == gridDim.x - 1) {
                *done_counter = 0;
                __threadfence();
                const int f = *reinterpret_cast<volatile int *>(rebuild_flag);
                *host_flag = (step_tag << 2) | (f & 3);
                __threadfence_system();
            }
        }
    }
    if (push_block) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();  // this block's peer stores are performed before its arrival is counted
            const unsigned n_push_blocks = gridDim.x - (unsigned)(n_interior / (int)blockDim.x);
            if (atomicAdd(hp.done_counter, 1u) == n_push_blocks - 1) {
                *hp.done_counter = 0;  // the next launch follows in stream order
                __threadfence_system();
                halo_st_release_sys(hp.sig_ready_prev, hp.epoch);
                halo_st_release_sys(hp.sig_ready_next, hp.epoch);
            }
        }
    }
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

// positions as packed float3: in slot order (orig == nullptr: a decomposed rank's owned block) or scattered to the caller's ids
__global__ void pack_xyz_kernel(int n, const float4 *__restrict__ sorted, const int *__restrict__ orig, float *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 v = sorted[k];
    const size_t o = 3 * (size_t)(orig ? orig[k] : k);
    out[o] = v.x; out[o + 1] = v.y; out[o + 2] = v.z;
}

__global__ void l2_flush_kernel(float4 *buf, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace

#ifdef MC_HAVE_LAUNCH  // the stand-ins of tests/cpp/shim/ and shim_mt/ have no launcher; shim_fiber/ has
void launch_kick_drift(int n_rows, float4 *xyzq, float4 *vel, const float4 *force, const float *ext_force,
                       const int *orig, const uint8_t *flags, const float4 *xref, float kick, float drift,
                       float max_disp, float lookahead, int *rebuild_flag, cudaStream_t st, int64_t *launches,
                       uint32_t *done_counter, int *host_flag, int step_tag) {
    if (n_rows <= 0) return;
    MC_LAUNCH(kick_drift_kernel<false>, div_up(n_rows, 256), 256, 0, st, n_rows, xyzq, vel, force, ext_force, orig, flags, xref, kick,
                                                                 drift, max_disp, lookahead, rebuild_flag, HaloPush{},
                                                                 done_counter, host_flag, step_tag);
    *launches += 1;
}

void launch_kick_drift_halo(int n_rows, float4 *xyzq, float4 *vel, const float4 *force, const float *ext_force,
                            const int *orig, const uint8_t *flags, const float4 *xref, float kick, float drift,
                            float max_disp, int *rebuild_flag, const HaloPush &hp, cudaStream_t st, int64_t *launches) {
    if (n_rows <= 0) return;
    MC_LAUNCH(kick_drift_kernel<true>, div_up(n_rows, 256), 256, 0, st, n_rows, xyzq, vel, force, ext_force, orig, flags, xref, kick,
                                                                drift, max_disp, 0.f, rebuild_flag, hp, nullptr, nullptr, 0);
    *launches += 1;
}

void launch_gather_to_orig(int n, const float4 *sorted, const int *orig, float4 *out, cudaStream_t st, int64_t *launches) {
    if (n <= 0) return;
    MC_LAUNCH(gather_to_orig_kernel, div_up(n, 256), 256, 0, st, n, sorted, orig, out);
    *launches += 1;
}

void launch_scatter_from_orig(int n, const float4 *in_orig, const int *orig, float4 *sorted, int keep_w, cudaStream_t st,
                              int64_t *launches) {
    if (n <= 0) return;
    MC_LAUNCH(scatter_from_orig_kernel, div_up(n, 256), 256, 0, st, n, in_orig, orig, sorted, keep_w);
    *launches += 1;
}

void launch_pack_xyz(int n, const float4 *sorted, const int *orig, float *out, cudaStream_t st, int64_t *launches) {
    if (n <= 0) return;
    MC_LAUNCH(pack_xyz_kernel, div_up(n, 256), 256, 0, st, n, sorted, orig, out);
    *launches += 1;
}

void launch_l2_flush(float4 *buf, size_t n_float4, cudaStream_t st, int64_t *launches) {
    MC_LAUNCH(l2_flush_kernel, 1184, 256, 0, st, buf, n_float4);
    *launches += 1;
}
#endif  // MC_HOST_SHIM
