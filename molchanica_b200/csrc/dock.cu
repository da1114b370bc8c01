// dock.cu -- docking pose-energy scan: for every rigid pose the R x L receptor-ligand pair sum of
// reference src/docking/legacy/mod.rs:210-383 (calc_binding_energy) with the weights of :174-200
// (BindingEnergy::new).  The reference fills an R x L distance cache per pose (:221-229) and sums
// it with rayon + AVX f32x8 (:235-262); here one CTA owns one pose, the posed ligand (L float4 +
// L meta words) and the T_rec x T_lig LJ table live in shared memory, each thread streams receptor
// atoms (coalesced float4, L2-resident: the 80 KB receptor is shared by all CTAs) and keeps the
// six partial sums in registers; a warp-shuffle + shared-memory tree finishes the pose.
//
// Bound: FP32 issue + MUFU (rcp/rsqrt), not HBM -- the receptor and table are on chip.
//
//   ligand atom a at pose p : x = anchor_p + R(q_p) (lig_a - lig_anchor)           (:149-158)
//   vdw          = sum 4 eps ((sigma/r)^12 - (sigma/r)^6)                          (:235-262; lj_V, cuda/util.cu:74-90)
//   hydrophobic  = sum over flagged pairs with r < 4.25 of -0.2 (1 - r/4.25)       (:305-321, :70)
//   electrostatic= | sum coulomb_force(rec -> lig) |, softening 1e-6               (:332-375, exact direct sum)
//   coulomb_e    = sum q_r q_l / r
//   score        = vdw + (-1.2) n_hbond + hydrophobic + 10 electrostatic, n_hbond = 0 (external crate)
#include "common.cuh"
#include "dock.cuh"

namespace {

constexpr int DOCK_THREADS = 128;
constexpr float HYDROPHOBIC_CUTOFF = 4.25f;

#ifndef MC_HOST_SHIM
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#else  // tests/cpp/dock_kernel_host.cpp runs this file's kernel on the CPU: no PTX there
inline float rcp_approx(float x) { return 1.0f / x; }
inline float rsqrt_approx(float x) { return 1.0f / sqrtf(x); }
#endif

__global__ void __launch_bounds__(DOCK_THREADS) dock_score_kernel(int n_rec, const float4 *__restrict__ rec,
                                                                   const uint32_t *__restrict__ rec_meta, int n_lig,
                                                                   const float4 *__restrict__ lig,
                                                                   const uint32_t *__restrict__ lig_meta, float3 anchor0,
                                                                   int n_rec_types, int n_lig_types,
                                                                   const float2 *__restrict__ ljtab,
                                                                   const float *__restrict__ poses,
                                                                   float *__restrict__ out) {
    MC_DYN_SHARED(float4, smem);
    float4 *lp = smem;                                             // n_lig posed atoms (x, y, z, q)
    uint32_t *lmeta = reinterpret_cast<uint32_t *>(lp + n_lig);    // n_lig meta words
    float2 *tab = reinterpret_cast<float2 *>(lmeta + ((n_lig + 3) & ~3));  // T_rec * T_lig
    const int pose = blockIdx.x;
    const float *ps = poses + 7 * (size_t)pose;
    {
        // The pose transform runs in f64 and is rounded once to f32, exactly as the reference does
        // (Pose{anchor_posit, orientation} are f64, lig_posits are Vec3F32; legacy/mod.rs:149-158,
        // :210-214).  Explicit _rn intrinsics (no fma contraction) keep the L posed points
        // bit-identical to the CPU path; it is ~40 atoms per pose, invisible next to R x L pairs.
        double qw = ps[3], qx = ps[4], qy = ps[5], qz = ps[6];
        const double qn = sqrt(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(qw, qw), __dmul_rn(qx, qx)), __dmul_rn(qy, qy)),
                                         __dmul_rn(qz, qz)));
        qw = __ddiv_rn(qw, qn); qx = __ddiv_rn(qx, qn); qy = __ddiv_rn(qy, qn); qz = __ddiv_rn(qz, qn);
        for (int a = threadIdx.x; a < n_lig; a += DOCK_THREADS) {
            const float4 l = lig[a];
            const double vx = __dsub_rn((double)l.x, (double)anchor0.x), vy = __dsub_rn((double)l.y, (double)anchor0.y),
                         vz = __dsub_rn((double)l.z, (double)anchor0.z);
            // v' = v + 2 (w (u x v) + u x (u x v))
            const double cx = __dsub_rn(__dmul_rn(qy, vz), __dmul_rn(qz, vy));
            const double cy = __dsub_rn(__dmul_rn(qz, vx), __dmul_rn(qx, vz));
            const double cz = __dsub_rn(__dmul_rn(qx, vy), __dmul_rn(qy, vx));
            const double dx = __dsub_rn(__dmul_rn(qy, cz), __dmul_rn(qz, cy));
            const double dy = __dsub_rn(__dmul_rn(qz, cx), __dmul_rn(qx, cz));
            const double dz = __dsub_rn(__dmul_rn(qx, cy), __dmul_rn(qy, cx));
            const double ox = __dadd_rn(vx, __dmul_rn(2.0, __dadd_rn(__dmul_rn(qw, cx), dx)));
            const double oy = __dadd_rn(vy, __dmul_rn(2.0, __dadd_rn(__dmul_rn(qw, cy), dy)));
            const double oz = __dadd_rn(vz, __dmul_rn(2.0, __dadd_rn(__dmul_rn(qw, cz), dz)));
            lp[a] = make_float4((float)__dadd_rn(ox, (double)ps[0]), (float)__dadd_rn(oy, (double)ps[1]),
                                (float)__dadd_rn(oz, (double)ps[2]), l.w);
            lmeta[a] = lig_meta[a];
        }
        for (int t = threadIdx.x; t < n_rec_types * n_lig_types; t += DOCK_THREADS) tab[t] = ljtab[t];
    }
    __syncthreads();

    float vdw = 0.f, hyd = 0.f, ec = 0.f, fx = 0.f, fy = 0.f, fz = 0.f;
    for (int r = threadIdx.x; r < n_rec; r += DOCK_THREADS) {
        const float4 xr = __ldg(rec + r);
        const uint32_t mr = __ldg(rec_meta + r);
        const float2 *row = tab + (mr & 0xffffu) * n_lig_types;
        const bool hr = (mr >> 16) & 1u;
#pragma unroll 4
        for (int a = 0; a < n_lig; ++a) {
            const float4 xl = lp[a];
            const uint32_t ml = lmeta[a];
            const float2 lj = row[ml & 0xffffu];
            const float dx = xl.x - xr.x, dy = xl.y - xr.y, dz = xl.z - xr.z;
            const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            const float ir2 = rcp_approx(r2);
            const float s2 = lj.x * ir2, s6 = s2 * s2 * s2;
            vdw = fmaf(lj.y * s6, s6 - 1.f, vdw);
            const float ir = rsqrt_approx(r2);
            const float qq = xr.w * xl.w;
            ec = fmaf(qq, ir, ec);
            const float fm = qq * ir * rcp_approx(r2 + MC_SOFTENING_SQ);
            fx = fmaf(dx, fm, fx); fy = fmaf(dy, fm, fy); fz = fmaf(dz, fm, fz);
            if (hr && ((ml >> 16) & 1u)) {
                const float rr = r2 * ir;
                if (rr < HYDROPHOBIC_CUTOFF) hyd += -0.2f * fmaxf(1.0f - rr * (1.0f / HYDROPHOBIC_CUTOFF), 0.f);
            }
        }
    }
    float v[6] = {vdw, hyd, ec, fx, fy, fz};
    __shared__ float red[6][DOCK_THREADS / 32];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(MC_FULL_MASK, v[k], d);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t[6];
        for (int k = 0; k < 6; ++k) {
            t[k] = 0.f;
            for (int w = 0; w < DOCK_THREADS / 32; ++w) t[k] += red[k][w];
        }
        const float es = sqrtf(t[3] * t[3] + t[4] * t[4] + t[5] * t[5]);
        float *o = out + 5 * (size_t)pose;
        o[0] = 1.f * t[0] + 0.f + 1.f * t[1] + 10.f * es;
        o[1] = t[0]; o[2] = t[1]; o[3] = es; o[4] = t[2];
    }
}

}  // namespace

#ifdef MC_HAVE_LAUNCH  // the stand-ins of tests/cpp/shim/ and shim_mt/ have no launcher; shim_fiber/ has
size_t dock_smem_bytes(int n_lig, int n_rec_types, int n_lig_types) {
    return sizeof(float4) * n_lig + sizeof(uint32_t) * ((n_lig + 3) & ~3) + sizeof(float2) * n_rec_types * n_lig_types;
}

cudaError_t dock_prepare() {
    return cudaFuncSetAttribute(dock_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

void launch_dock_score(int n_rec, const float4 *rec, const uint32_t *rec_meta, int n_lig, const float4 *lig,
                       const uint32_t *lig_meta, float3 lig_anchor, int n_rec_types, int n_lig_types,
                       const float2 *ljtab, int n_poses, const float *poses, float *out, cudaStream_t st,
                       int64_t *launches) {
    if (n_poses <= 0) return;
    MC_LAUNCH(dock_score_kernel, n_poses, DOCK_THREADS, dock_smem_bytes(n_lig, n_rec_types, n_lig_types), st, 
        n_rec, rec, rec_meta, n_lig, lig, lig_meta, lig_anchor, n_rec_types, n_lig_types, ljtab, poses, out);
    *launches += 1;
}
#endif  // MC_HOST_SHIM
