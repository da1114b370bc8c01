// md_fused.cu -- many MD steps in ONE launch for small systems ("CUDA Graphs / persistent multi-step kernel";
// the reference's GUI calls MdState::step ten times per frame, src/md/mod.rs:45,737-749).  At a thousand atoms a step is a
// handful of microseconds of arithmetic; run as separate launches (kick+drift, pair forces, 1-4 pairs, bonded terms) plus a
// host poll of the rebuild flag per step it measured 40 us per step on C2 -- launch latency, not work.  Here a persistent
// cooperative grid keeps the whole system on chip between steps and replaces kernel boundaries by grid-wide barriers:
//
//     repeat n_steps:   [kick + drift, displacement check]  --grid barrier--  stop if an atom outran skin/2 (or blew up)
//                       [pair forces over the Verlet rows, each row's 1-4 partners]  (--grid barrier-- [bonded terms])
//                       --grid barrier--
//     closing half kick
//
// Same arithmetic as the per-launch path, term for term: kick / drift of integrate.cu, pair_term of pair_terms.cuh in
// the same lane layout and order as pair_force.cu (8 lanes per row, two entries in flight), the 1-4 rows of
// pair_force.cu, bonded_term_apply of bonded_device.cuh.  The displacement criterion is the synchronous one (no
// look-ahead): the step that trips it ends after its drift, the host rebuilds the list, evaluates the forces and
// relaunches for the remaining steps.  Plain single-GPU NVE systems only (no constraints, thermostat, barostat, SPME,
// virtual sites, centre-of-mass removal); everything else takes the per-launch path.
#include <algorithm>

#include "common.cuh"
#include "md_fused.cuh"

#ifndef MC_HOST_SHIM  // the stand-ins of tests/cpp/ have no cooperative launch: the per-launch path covers these systems there
#include <cooperative_groups.h>

#include "bonded_device.cuh"
#include "pair_terms.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int FUSED_THREADS = 256;
constexpr int FUSED_LANES = 8;

template <bool MULTI, int COUL, bool WRAP>
__device__ __forceinline__ void fused_row(const float4 xi, const uint32_t *__restrict__ lst, uint32_t cnt, int sub, const float4 *xyzq,
                                          const uint16_t *__restrict__ type, const float2 *row, const NbParams &p, bool lj_on, Acc &a) {
    const float2 lj1 = make_float2(p.sig2, p.eps24);
    const float rc2_lj = lj_on ? p.rc2_lj : -1.f;
    uint32_t k = sub;
    // positions are written by this very kernel (previous phase): ordinary loads, not the read-only path
    for (; k + FUSED_LANES < cnt; k += 2 * FUSED_LANES) {
        const uint32_t j0 = __ldg(lst + k), j1 = __ldg(lst + k + FUSED_LANES);
        const float4 x0 = xyzq[j0], x1 = xyzq[j1];
        float2 l0 = lj1, l1 = lj1;
        if (MULTI) { l0 = row[__ldg(type + j0)]; l1 = row[__ldg(type + j1)]; }
        pair_term<COUL, WRAP, false>(xi, x0, l0, p, rc2_lj, a);
        pair_term<COUL, WRAP, false>(xi, x1, l1, p, rc2_lj, a);
    }
    if (k < cnt) {
        const uint32_t j0 = __ldg(lst + k);
        const float4 x0 = xyzq[j0];
        float2 l0 = lj1;
        if (MULTI) l0 = row[__ldg(type + j0)];
        pair_term<COUL, WRAP, false>(xi, x0, l0, p, rc2_lj, a);
    }
}

template <bool MULTI, int COUL, bool PBC>
__global__ void __launch_bounds__(FUSED_THREADS) md_fused_kernel(const FusedArgs A) {
    cg::grid_group grid = cg::this_grid();
    MC_DYN_SHARED(float2, s_tab);
    if (MULTI) {
        for (int t = threadIdx.x; t < A.p.n_types * A.p.n_types; t += blockDim.x) s_tab[t] = A.ljtab[t];
        __syncthreads();
    }
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
    const int sub = threadIdx.x % FUSED_LANES;
    int s = 0, flag = 0;
    for (; s < A.n_steps; ++s) {
        // ---- kick + drift (integrate.cu), displacement criterion without look-ahead
        const float kick = (s == 0 && A.first_half) ? 0.5f * A.dt : A.dt;
        int my = 0;
        for (int i = tid; i < A.n; i += n_threads) {
            if (A.flags[i] & MC_FLAG_STATIC) continue;
            float4 v = A.vel[i];
            const float4 f = A.force[i];
            float fx = f.x, fy = f.y, fz = f.z;
            if (A.ext_force) {
                const int o = A.orig[i];
                fx += A.ext_force[3 * o]; fy += A.ext_force[3 * o + 1]; fz += A.ext_force[3 * o + 2];
            }
            const float sc = v.w * kick * MC_ACCEL_CONV;
            v.x = fmaf(fx, sc, v.x); v.y = fmaf(fy, sc, v.y); v.z = fmaf(fz, sc, v.z);
            A.vel[i] = v;
            float4 x = A.xyzq[i];
            x.x = fmaf(v.x, A.dt, x.x); x.y = fmaf(v.y, A.dt, x.y); x.z = fmaf(v.z, A.dt, x.z);
            A.xyzq[i] = x;
            const float4 r = A.xref[i];
            const float dx = x.x - r.x, dy = x.y - r.y, dz = x.z - r.z;
            if (dx * dx + dy * dy + dz * dz > A.max_disp * A.max_disp) my |= 1;
            if (!(fabsf(x.x) + fabsf(x.y) + fabsf(x.z) < 1.0e30f)) my |= 2;
        }
        if (my) atomicOr(A.rebuild_flag, my);
        grid.sync();
        flag = *reinterpret_cast<volatile int *>(A.rebuild_flag);
        if (flag & 3) break;  // grid-uniform: this step ends after its drift; the host rebuilds / reports
        // ---- pair forces over the Verlet rows (pair_force.cu's layout), the row's 1-4 partners added by its first lane
        // (whole warps enter every iteration -- the reduction below shuffles with the full mask; rows past the end are dead lanes)
        for (int rb = (tid / 32) * (32 / FUSED_LANES); rb < A.n; rb += n_threads / FUSED_LANES) {
            const int r_raw = rb + (threadIdx.x & 31) / FUSED_LANES;
            const bool live = r_raw < A.n;
            const int r = live ? r_raw : A.n - 1;
            const float4 xi = A.xyzq[r];
            const uint32_t start = __ldg(A.nbr_start + r), cnt = live ? __ldg(A.nbr_count + r) : 0u;
            const int ti = MULTI ? (int)__ldg(A.type + r) : 0;
            const float2 *row = MULTI ? s_tab + ti * A.p.n_types : nullptr;
            const bool wrap = PBC && !(__ldg(A.flags + r) & MC_FLAG_INTERIOR);
            Acc a = {0.f, 0.f, 0.f, 0.f};
            if (wrap) fused_row<MULTI, COUL, true>(xi, A.nbr_list + start, cnt, sub, A.xyzq, A.type, row, A.p, A.lj_on != 0, a);
            else fused_row<MULTI, COUL, false>(xi, A.nbr_list + start, cnt, sub, A.xyzq, A.type, row, A.p, A.lj_on != 0, a);
#pragma unroll
            for (int d = FUSED_LANES / 2; d > 0; d >>= 1) {
                a.fx += __shfl_xor_sync(MC_FULL_MASK, a.fx, d);
                a.fy += __shfl_xor_sync(MC_FULL_MASK, a.fy, d);
                a.fz += __shfl_xor_sync(MC_FULL_MASK, a.fz, d);
            }
            if (live && sub == 0) {
                if (A.p14_start) {  // Amber 1-4 rows (pairs14_kernel of pair_force.cu): no cutoff, scaled LJ / Coulomb
                    const int oi = A.orig[r];
                    for (int e = A.p14_start[oi]; e < A.p14_start[oi + 1]; ++e) {
                        const int j = A.slot_of_orig[A.p14_idx[e]];
                        if (j < 0) continue;
                        const float4 xj = A.xyzq[j];
                        float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                        if (A.p.periodic) {
                            dx -= rintf(dx * A.p.inv_ext[0]) * A.p.ext[0];
                            dy -= rintf(dy * A.p.inv_ext[1]) * A.p.ext[1];
                            dz -= rintf(dz * A.p.inv_ext[2]) * A.p.ext[2];
                        }
                        const float r2 = dx * dx + dy * dy + dz * dz;
                        const float ir2 = 1.0f / r2;
                        float ff = 0.f;
                        if (A.lj_on) {
                            const float2 lj = A.ljtab[ti * A.p.n_types + (MULTI ? (int)A.type[j] : 0)];
                            const float s2 = lj.x * ir2, s6 = s2 * s2 * s2;
                            ff += A.s14_lj * lj.y * s6 * (2.f * s6 - 1.f) * ir2;
                        }
                        if (COUL != MC_COULOMB_NONE) {
                            const float qq = xi.w * xj.w, ir = rsqrtf(r2);
                            ff += A.s14_q * qq * ir / (r2 + MC_SOFTENING_SQ);
                        }
                        a.fx += dx * ff; a.fy += dy * ff; a.fz += dz * ff;
                    }
                }
                A.force[r] = make_float4(a.fx, a.fy, a.fz, 0.f);
            }
        }
        const int n_bonded = A.bt.n_bonds + A.bt.n_angles + A.bt.n_dihedrals;
        if (n_bonded > 0) {
            grid.sync();  // every row is written before the bonded terms add to it
            for (int t = tid; t < n_bonded; t += n_threads) {
                float e, w;
                int kind;
                bonded_term_apply(t, A.bt, A.slot_of_orig, A.xyzq, A.p, A.force, e, w, kind);
            }
        }
        grid.sync();
    }
    if (s == A.n_steps) {
        // closing half kick of the last step
        for (int i = tid; i < A.n; i += n_threads) {
            if (A.flags[i] & MC_FLAG_STATIC) continue;
            float4 v = A.vel[i];
            const float4 f = A.force[i];
            float fx = f.x, fy = f.y, fz = f.z;
            if (A.ext_force) {
                const int o = A.orig[i];
                fx += A.ext_force[3 * o]; fy += A.ext_force[3 * o + 1]; fz += A.ext_force[3 * o + 2];
            }
            const float sc = v.w * 0.5f * A.dt * MC_ACCEL_CONV;
            v.x = fmaf(fx, sc, v.x); v.y = fmaf(fy, sc, v.y); v.z = fmaf(fz, sc, v.z);
            A.vel[i] = v;
        }
    }
    if (tid == 0) {
        A.out[0] = s == A.n_steps ? A.n_steps : s + 1;  // drifts completed; < n_steps or flagged: forces of the last one pending
        A.out[1] = flag;
    }
}

}  // namespace

// Largest system the fused kernel takes: the grid must be co-resident (cooperative launch) and the win is launch latency
int md_fused_max_atoms() { return 65536; }

cudaError_t md_fused_prepare() {
    cudaError_t e = cudaSuccess;
#define MC_FA(M, C, P) if (e == cudaSuccess) e = cudaFuncSetAttribute(md_fused_kernel<M, C, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
#define MC_FA_C(M) MC_FA(M, MC_COULOMB_NONE, true) MC_FA(M, MC_COULOMB_NONE, false) MC_FA(M, MC_COULOMB_PLAIN, true) MC_FA(M, MC_COULOMB_PLAIN, false) \
    MC_FA(M, MC_COULOMB_ERFC, true) MC_FA(M, MC_COULOMB_ERFC, false)
    MC_FA_C(true) MC_FA_C(false)
#undef MC_FA_C
#undef MC_FA
    return e;
}

template <bool MULTI, int COUL, bool PBC>
static cudaError_t launch_fused_t(const FusedArgs &A, int n_sms, cudaStream_t st) {
    const size_t smem = MULTI ? sizeof(float2) * A.p.n_types * A.p.n_types : 0;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, md_fused_kernel<MULTI, COUL, PBC>, FUSED_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    const int want = (int)div_up((size_t)A.n * FUSED_LANES, FUSED_THREADS);
    // a grid barrier costs more the more blocks take part: no more blocks than the rows need, at most two per SM
    const int grid = std::max(1, std::min(want, n_sms * std::min(per_sm, 2)));
    FusedArgs a = A;
    void *args[] = {&a};
    return cudaLaunchCooperativeKernel(reinterpret_cast<void *>(md_fused_kernel<MULTI, COUL, PBC>), dim3((unsigned)grid), dim3(FUSED_THREADS), args,
                                       smem, st);
}

cudaError_t launch_md_fused(const FusedArgs &A, bool multi, int coul, bool pbc, int n_sms, cudaStream_t st, int64_t *launches) {
    cudaError_t e;
#define MC_FU_P(M, C) (pbc ? launch_fused_t<M, C, true>(A, n_sms, st) : launch_fused_t<M, C, false>(A, n_sms, st))
#define MC_FU_C(M) (coul == MC_COULOMB_NONE ? MC_FU_P(M, MC_COULOMB_NONE) : (coul == MC_COULOMB_PLAIN ? MC_FU_P(M, MC_COULOMB_PLAIN) : MC_FU_P(M, MC_COULOMB_ERFC)))
    e = multi ? MC_FU_C(true) : MC_FU_C(false);
#undef MC_FU_C
#undef MC_FU_P
    *launches += 1;
    return e;
}
#elif defined(MC_HAVE_LAUNCH)
int md_fused_max_atoms() { return 0; }
cudaError_t md_fused_prepare() { return cudaSuccess; }
cudaError_t launch_md_fused(const FusedArgs &, bool, int, bool, int, cudaStream_t, int64_t *) { return cudaErrorNotSupported; }
#endif
