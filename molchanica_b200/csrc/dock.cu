// dock.cu -- docking pose-energy scan: for every rigid pose the R x L receptor-ligand pair sum of
// reference src/docking/legacy/mod.rs:210-383 (calc_binding_energy) with the weights of :174-200
// (BindingEnergy::new).  The reference fills an R x L distance cache per pose (:221-229) and sums
// it with rayon + AVX f32x8 (:235-262); here one CTA owns one pose, the posed ligand (L float4 +
// L meta words) and the T_rec x T_lig LJ table live in shared memory, each thread streams receptor
// atoms (coalesced float4, L2-resident: the 80 KB receptor is shared by all CTAs) and keeps the
// six partial sums in registers; a warp-shuffle + shared-memory tree finishes the pose.
//
// Bound: FP32 issue, not HBM -- the receptor and table are on chip.  The inner loop takes TWO ligand atoms per
// iteration in packed fp32 (f32x2: the posed ligand is kept as (a, a+1) pairs per component, the LJ parameters as pairs
// per receptor type, so every LDS.64 delivers a packed operand) and ONE MUFU per pair: 1/r from rsqrt, 1/r^2 = (1/r)^2,
// and the softened 1/(r^2 + 1e-6) = (1/r^2)(1 - 1e-6/r^2 + ...) to first order (exact to 1e-8 for r > 0.1 A).
//
//   ligand atom a at pose p : x = anchor_p + R(q_p) (lig_a - lig_anchor)           (:149-158)
//   vdw          = sum 4 eps ((sigma/r)^12 - (sigma/r)^6)                          (:235-262; lj_V, cuda/util.cu:74-90)
//   hydrophobic  = sum over flagged pairs with r < 4.25 of -0.2 (1 - r/4.25)       (:305-321, :70)
//   electrostatic= | sum coulomb_force(rec -> lig) |, softening 1e-6               (:332-375, exact direct sum)
//   coulomb_e    = sum q_r q_l / r
//   score        = vdw + (-1.2) n_hbond + hydrophobic + 10 electrostatic, n_hbond = 0 (external crate)
#include "common.cuh"
#include "dock.cuh"
#include "pair_terms.cuh"  // rsqrt_approx, packed fp32 helpers

namespace {

constexpr int DOCK_THREADS = 128;
constexpr float HYDROPHOBIC_CUTOFF = 4.25f;

__global__ void __launch_bounds__(DOCK_THREADS) dock_score_kernel(int n_rec, const float4 *__restrict__ rec,
                                                                   const uint32_t *__restrict__ rec_meta, int n_lig,
                                                                   const float4 *__restrict__ lig,
                                                                   const uint32_t *__restrict__ lig_meta, float3 anchor0,
                                                                   int n_rec_types, int n_lig_types,
                                                                   const float2 *__restrict__ ljtab,
                                                                   const float *__restrict__ poses, int pose_stride, int n_flex,
                                                                   const int2 *__restrict__ flex_axis,
                                                                   const uint8_t *__restrict__ flex_mask,
                                                                   float *__restrict__ out) {
    MC_DYN_SHARED(float2, smem);
    // posed ligand as pairs of atoms (a, a+1) per component; an odd count is padded with a far, neutral, sigma = 0 atom
    const int np = (n_lig + 1) >> 1;
    float2 *lx2 = smem, *ly2 = lx2 + np, *lz2 = ly2 + np, *lq2 = lz2 + np;
    float2 *s2tab = lq2 + np;                    // [n_rec_types][np] (sigma^2 of the pair's two atoms)
    float2 *e4tab = s2tab + n_rec_types * np;    // [n_rec_types][np] (4 eps)
    uint32_t *lhyd = reinterpret_cast<uint32_t *>(e4tab + n_rec_types * np);  // [np] hydrophobic flags, bit 0 / bit 1
    const int pose = blockIdx.x;
    const float *ps = poses + (size_t)pose_stride * (size_t)pose;
    // flexible ligand: the conformer of this pose in f64 (3 n_lig doubles behind the tables)
    double *conf = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(lhyd + np) + 7) & ~(uintptr_t)7);
    if (n_flex > 0) {
        for (int a = threadIdx.x; a < n_lig; a += DOCK_THREADS) {
            const float4 l = lig[a];
            conf[3 * a] = (double)l.x; conf[3 * a + 1] = (double)l.y; conf[3 * a + 2] = (double)l.z;
        }
        for (int f = 0; f < n_flex; ++f) {
            __syncthreads();
            const int2 ax = flex_axis[f];
            const double p0x = conf[3 * ax.x], p0y = conf[3 * ax.x + 1], p0z = conf[3 * ax.x + 2];
            const double p1x = conf[3 * ax.y], p1y = conf[3 * ax.y + 1], p1z = conf[3 * ax.y + 2];
            __syncthreads();  // every thread has read the axis before anybody moves an atom
            double ux = p1x - p0x, uy = p1y - p0y, uz = p1z - p0z;
            const double un = 1.0 / sqrt(ux * ux + uy * uy + uz * uz);
            ux *= un; uy *= un; uz *= un;
            double sn, cs;
            sincos((double)ps[7 + f], &sn, &cs);
            for (int a = threadIdx.x; a < n_lig; a += DOCK_THREADS) {
                if (!flex_mask[(size_t)f * n_lig + a]) continue;
                // Rodrigues: v' = v cos t + (u x v) sin t + u (u . v)(1 - cos t), about the axis through a1
                const double vx = conf[3 * a] - p1x, vy = conf[3 * a + 1] - p1y, vz = conf[3 * a + 2] - p1z;
                const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
                const double dt = (ux * vx + uy * vy + uz * vz) * (1.0 - cs);
                conf[3 * a] = p1x + vx * cs + cx * sn + ux * dt;
                conf[3 * a + 1] = p1y + vy * cs + cy * sn + uy * dt;
                conf[3 * a + 2] = p1z + vz * cs + cz * sn + uz * dt;
            }
        }
        __syncthreads();
    }
    {
        // The pose transform runs in f64 and is rounded once to f32, exactly as the reference does
        // (Pose{anchor_posit, orientation} are f64, lig_posits are Vec3F32; legacy/mod.rs:149-158,
        // :210-214).  Explicit _rn intrinsics (no fma contraction) keep the L posed points
        // bit-identical to the CPU path; it is ~40 atoms per pose, invisible next to R x L pairs.
        double qw = ps[3], qx = ps[4], qy = ps[5], qz = ps[6];
        const double qn = sqrt(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(qw, qw), __dmul_rn(qx, qx)), __dmul_rn(qy, qy)),
                                         __dmul_rn(qz, qz)));
        qw = __ddiv_rn(qw, qn); qx = __ddiv_rn(qx, qn); qy = __ddiv_rn(qy, qn); qz = __ddiv_rn(qz, qn);
        float *lxs = reinterpret_cast<float *>(lx2), *lys = reinterpret_cast<float *>(ly2), *lzs = reinterpret_cast<float *>(lz2),
              *lqs = reinterpret_cast<float *>(lq2);
        for (int a = threadIdx.x; a < 2 * np; a += DOCK_THREADS) {
            if (a >= n_lig) { lxs[a] = lys[a] = lzs[a] = 1.0e6f; lqs[a] = 0.f; continue; }
            const float4 l = lig[a];
            const double lx = n_flex > 0 ? conf[3 * a] : (double)l.x, ly = n_flex > 0 ? conf[3 * a + 1] : (double)l.y,
                         lz = n_flex > 0 ? conf[3 * a + 2] : (double)l.z;
            const double vx = __dsub_rn(lx, (double)anchor0.x), vy = __dsub_rn(ly, (double)anchor0.y),
                         vz = __dsub_rn(lz, (double)anchor0.z);
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
            lxs[a] = (float)__dadd_rn(ox, (double)ps[0]);
            lys[a] = (float)__dadd_rn(oy, (double)ps[1]);
            lzs[a] = (float)__dadd_rn(oz, (double)ps[2]);
            lqs[a] = l.w;
        }
        // LJ parameters of every (receptor type, ligand atom) as pairs, hydrophobic flags of the ligand pairs
        float *s2s = reinterpret_cast<float *>(s2tab), *e4s = reinterpret_cast<float *>(e4tab);
        for (int t = threadIdx.x; t < n_rec_types * 2 * np; t += DOCK_THREADS) {
            const int rt = t / (2 * np), a = t - rt * 2 * np;
            float2 lj = make_float2(0.f, 0.f);
            if (a < n_lig) lj = ljtab[rt * n_lig_types + (int)(lig_meta[a] & 0xffffu)];
            s2s[t] = lj.x;
            e4s[t] = lj.y;
        }
        for (int p = threadIdx.x; p < np; p += DOCK_THREADS) {
            const uint32_t h0 = (lig_meta[2 * p] >> 16) & 1u, h1 = (2 * p + 1 < n_lig) ? ((lig_meta[2 * p + 1] >> 16) & 1u) : 0u;
            lhyd[p] = h0 | (h1 << 1);
        }
    }
    __syncthreads();

    float hyd = 0.f;
    float2 vdw2 = make_float2(0.f, 0.f), ec2 = vdw2, fx2 = vdw2, fy2 = vdw2, fz2 = vdw2;
    const float2 one_n = make_float2(-1.f, -1.f), one = make_float2(1.f, 1.f), soft_n = make_float2(-MC_SOFTENING_SQ, -MC_SOFTENING_SQ);
    for (int r = threadIdx.x; r < n_rec; r += DOCK_THREADS) {
        const float4 xr = __ldg(rec + r);
        const uint32_t mr = __ldg(rec_meta + r);
        const float2 *s2row = s2tab + (mr & 0xffffu) * np, *e4row = e4tab + (mr & 0xffffu) * np;
        const bool hr = (mr >> 16) & 1u;
        const float2 nx = make_float2(-xr.x, -xr.x), ny = make_float2(-xr.y, -xr.y), nz = make_float2(-xr.z, -xr.z),
                     qr = make_float2(xr.w, xr.w);
#pragma unroll 2
        for (int p = 0; p < np; ++p) {
            const float2 dx = mc_add2(lx2[p], nx), dy = mc_add2(ly2[p], ny), dz = mc_add2(lz2[p], nz);
            const float2 r2 = mc_fma2(dx, dx, mc_fma2(dy, dy, mc_mul2(dz, dz)));
            const float2 ir = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
            const float2 ir2 = mc_mul2(ir, ir);
            const float2 s2 = mc_mul2(s2row[p], ir2);
            const float2 s6 = mc_mul2(mc_mul2(s2, s2), s2);
            vdw2 = mc_fma2(mc_mul2(e4row[p], s6), mc_add2(s6, one_n), vdw2);
            const float2 qq = mc_mul2(qr, lq2[p]);
            ec2 = mc_fma2(qq, ir, ec2);
            // qq / r / (r^2 + 1e-6) = qq (1/r)^3 (1 - 1e-6 / r^2) to first order
            const float2 fm = mc_mul2(mc_mul2(qq, mc_mul2(ir, ir2)), mc_fma2(soft_n, ir2, one));
            fx2 = mc_fma2(dx, fm, fx2); fy2 = mc_fma2(dy, fm, fy2); fz2 = mc_fma2(dz, fm, fz2);
            if (hr) {
                const uint32_t lh = lhyd[p];
                if (lh) {
                    const float2 rr = mc_mul2(r2, ir);
                    if ((lh & 1u) && rr.x < HYDROPHOBIC_CUTOFF) hyd += -0.2f * fmaxf(1.0f - rr.x * (1.0f / HYDROPHOBIC_CUTOFF), 0.f);
                    if ((lh & 2u) && rr.y < HYDROPHOBIC_CUTOFF) hyd += -0.2f * fmaxf(1.0f - rr.y * (1.0f / HYDROPHOBIC_CUTOFF), 0.f);
                }
            }
        }
    }
    const float vdw = vdw2.x + vdw2.y, ec = ec2.x + ec2.y, fx = fx2.x + fx2.y, fy = fy2.x + fy2.y, fz = fz2.x + fz2.y;
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
    (void)n_lig_types;
    const size_t np = (size_t)(n_lig + 1) / 2;
    // pairs of ligand atoms, LJ pairs per receptor type, hydrophobic flags, and the f64 conformer of a flexible ligand
    return sizeof(float2) * np * (4 + 2 * (size_t)n_rec_types) + sizeof(uint32_t) * np + 8 + sizeof(double) * 3 * (size_t)n_lig;
}

cudaError_t dock_prepare() {
    return cudaFuncSetAttribute(dock_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

void launch_dock_score(int n_rec, const float4 *rec, const uint32_t *rec_meta, int n_lig, const float4 *lig,
                       const uint32_t *lig_meta, float3 lig_anchor, int n_rec_types, int n_lig_types,
                       const float2 *ljtab, int n_poses, const float *poses, int pose_stride, int n_flex, const int2 *flex_axis,
                       const uint8_t *flex_mask, float *out, cudaStream_t st, int64_t *launches) {
    if (n_poses <= 0) return;
    MC_LAUNCH(dock_score_kernel, n_poses, DOCK_THREADS, dock_smem_bytes(n_lig, n_rec_types, n_lig_types), st, 
        n_rec, rec, rec_meta, n_lig, lig, lig_meta, lig_anchor, n_rec_types, n_lig_types, ljtab, poses, pose_stride, n_flex, flex_axis,
        flex_mask, out);
    *launches += 1;
}
#endif  // MC_HOST_SHIM
