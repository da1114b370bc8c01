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

__device__ __forceinline__ float fused_min_image_exact(float d, float ext, float inv_ext) {  // tile_build.cu's min_image_exact
    float q = __fmul_rn(d, inv_ext);
    float n = rintf(q);
    if (fabsf(q - n) > 0.4999f) n = rintf(__fdiv_rn(d, ext));
    return __fmaf_rn(-n, ext, d);
}

// RO: the rows are the host's list (read-only for the whole launch: non-coherent loads); else the kernel's own list
template <bool MULTI, int COUL, bool WRAP, bool RO>
__device__ __forceinline__ void fused_row(const float4 xi, const uint32_t *lst, uint32_t cnt, uint32_t sub, uint32_t lanes, const float4 *xyzq,
                                          const uint16_t *__restrict__ type, const float2 *row, const NbParams &p, bool lj_on, Acc &a) {
    const float2 lj1 = make_float2(p.sig2, p.eps24);
    const float rc2_lj = lj_on ? p.rc2_lj : -1.f;
    uint32_t k = sub;
    // positions are written by this very kernel (previous phase): ordinary loads, not the read-only path
    for (; k + lanes < cnt; k += 2 * lanes) {
        const uint32_t j0 = RO ? __ldg(lst + k) : lst[k], j1 = RO ? __ldg(lst + k + lanes) : lst[k + lanes];
        const float4 x0 = xyzq[j0], x1 = xyzq[j1];
        float2 l0 = lj1, l1 = lj1;
        if (MULTI) { l0 = row[__ldg(type + j0)]; l1 = row[__ldg(type + j1)]; }
        pair_term<COUL, WRAP, false>(xi, x0, l0, p, rc2_lj, a);
        pair_term<COUL, WRAP, false>(xi, x1, l1, p, rc2_lj, a);
    }
    if (k < cnt) {
        const uint32_t j0 = RO ? __ldg(lst + k) : lst[k];
        const float4 x0 = xyzq[j0];
        float2 l0 = lj1;
        if (MULTI) l0 = row[__ldg(type + j0)];
        pair_term<COUL, WRAP, false>(xi, x0, l0, p, rc2_lj, a);
    }
}

// Brute-force mode: the whole system's positions and types sit in shared memory (staged by every block after each drift), so
// the only global loads of a row are its indices -- and those are issued four entries per lane ahead of the arithmetic.  The
// force phase of a thousand-atom system is otherwise nothing but a chain of dependent L2 round trips (index -> position ->
// type, ~20 of them per row: measured 19 us per step at 1,231 atoms against ~2 us of arithmetic).  Per lane the entries are
// visited in the order of fused_row (k, k + lanes, k + 2 lanes, ...): same sums.
template <bool MULTI, int COUL, bool WRAP>
__device__ __forceinline__ void fused_row_staged(const float4 xi, const uint32_t *lst, uint32_t cnt, uint32_t sub, uint32_t lanes,
                                                 const float4 *s_pos, const uint16_t *s_type, const float2 *row, const NbParams &p,
                                                 bool lj_on, Acc &a) {
    const float2 lj1 = make_float2(p.sig2, p.eps24);
    const float rc2_lj = lj_on ? p.rc2_lj : -1.f;
    for (uint32_t k = sub; k < cnt; k += 4 * lanes) {
        uint32_t j[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) j[u] = (k + u * lanes < cnt) ? lst[k + u * lanes] : 0xffffffffu;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (j[u] != 0xffffffffu) {
                const float4 xj = s_pos[j[u]];
                float2 lj = lj1;
                if (MULTI) lj = row[s_type[j[u]]];
                pair_term<COUL, WRAP, false>(xi, xj, lj, p, rc2_lj, a);
            }
        }
    }
}

// kick + drift of integrate.cu over all atoms; displacement against `xref` and the non-finite guard go into *flag_word
__device__ __forceinline__ void fused_kick_drift(const FusedArgs &A, float kick, const float4 *xref, int *flag_word, int tid, int n_threads) {
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
        const float4 r = xref[i];
        const float dx = x.x - r.x, dy = x.y - r.y, dz = x.z - r.z;
        if (dx * dx + dy * dy + dz * dz > A.max_disp * A.max_disp) my |= 1;
        if (!(fabsf(x.x) + fabsf(x.y) + fabsf(x.z) < 1.0e30f)) my |= 2;
    }
    if (my) atomicOr(flag_word, my);
}

// closing half kick of the last step
__device__ __forceinline__ void fused_half_kick(const FusedArgs &A, int tid, int n_threads) {
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

// pair forces over the Verlet rows (pair_force.cu's layout; brute mode: the private fixed-stride rows), the row's 1-4
// partners added by its first lane.  Whole warps enter every iteration -- the reduction shuffles with the full mask; rows past
// the end are dead lanes.
template <bool MULTI, int COUL, bool PBC>
__device__ __forceinline__ void fused_forces(const FusedArgs &A, const float2 *s_tab, const float4 *s_pos, const uint16_t *s_type, int tid,
                                             int n_threads) {
    const int lanes = A.lanes;  // 8, 16 or 32 lanes per row: small systems spread a row over more lanes to fill the chip
    const uint32_t sub = threadIdx.x % lanes;
    for (int rb = (tid / 32) * (32 / lanes); rb < A.n; rb += n_threads / lanes) {
        const int r_raw = rb + (threadIdx.x & 31) / lanes;
        const bool live = r_raw < A.n;
        const int r = live ? r_raw : A.n - 1;
        const float4 xi = A.brute ? s_pos[r] : A.xyzq[r];
        uint32_t start, cnt;
        const uint32_t *list;
        if (A.brute) {  // (written by this kernel: plain loads)
            start = (uint32_t)r * A.bl_stride;
            cnt = live ? A.bl_count[r] : 0u;
            list = A.bl_list;
        } else {
            start = __ldg(A.nbr_start + r);
            cnt = live ? __ldg(A.nbr_count + r) : 0u;
            list = A.nbr_list;
        }
        const int ti = MULTI ? (A.brute ? (int)s_type[r] : (int)__ldg(A.type + r)) : 0;
        const float2 *row = MULTI ? s_tab + ti * A.p.n_types : nullptr;
        // (brute mode: atoms are never re-sorted, so the interior flag of the host's last build says nothing about where an
        // atom is now -- every row takes the minimum image)
        const bool wrap = PBC && (A.brute || !(__ldg(A.flags + r) & MC_FLAG_INTERIOR));
        Acc a = {0.f, 0.f, 0.f, 0.f};
        if (A.brute) {
            if (wrap) fused_row_staged<MULTI, COUL, true>(xi, list + start, cnt, sub, (uint32_t)lanes, s_pos, s_type, row, A.p, A.lj_on != 0, a);
            else fused_row_staged<MULTI, COUL, false>(xi, list + start, cnt, sub, (uint32_t)lanes, s_pos, s_type, row, A.p, A.lj_on != 0, a);
        } else {
            if (wrap) fused_row<MULTI, COUL, true, true>(xi, list + start, cnt, sub, (uint32_t)lanes, A.xyzq, A.type, row, A.p, A.lj_on != 0, a);
            else fused_row<MULTI, COUL, false, true>(xi, list + start, cnt, sub, (uint32_t)lanes, A.xyzq, A.type, row, A.p, A.lj_on != 0, a);
        }
        if (A.p14_start && live) {
            // Amber 1-4 partners of the row (pairs14_kernel of pair_force.cu: no cutoff, scaled LJ / Coulomb), dealt over the
            // row's lanes like the listed pairs: a single lane walking them one dependent gather after the other made the
            // atom with the most 1-4 partners the critical path of every step
            const int oi = A.orig[r];
            const int e1 = A.p14_start[oi + 1];
            for (int e = A.p14_start[oi] + (int)sub; e < e1; e += lanes) {
                const int j = A.slot_of_orig[A.p14_idx[e]];
                if (j < 0) continue;
                const float4 xj = A.brute ? s_pos[j] : A.xyzq[j];
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
                    const float2 lj = MULTI ? row[A.brute ? s_type[j] : A.type[j]] : make_float2(A.p.sig2, A.p.eps24);
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
        for (int d = lanes / 2; d > 0; d >>= 1) {
            a.fx += __shfl_xor_sync(MC_FULL_MASK, a.fx, d);
            a.fy += __shfl_xor_sync(MC_FULL_MASK, a.fy, d);
            a.fz += __shfl_xor_sync(MC_FULL_MASK, a.fz, d);
        }
        if (live && sub == 0) {
            A.force[r] = make_float4(a.fx, a.fy, a.fz, 0.f);
        }
    }
}

// Brute-force mode: the private list.  One warp per row, lanes sweep ALL atoms (lane = consecutive candidates), the warp-wide
// accept decision is one ballot and a hit's place in the row its rank in the ballot -- rows come out in ascending slot
// order.  The accept expression is the oracle's fp32 one (neighbor.cu / tile_build.cu): exact minimum image in a periodic
// box, ((dx*dx)+(dy*dy))+(dz*dz) < r_list^2 without contraction; excluded partners (original ids) never enter.
template <bool PBC>
__device__ __forceinline__ void fused_build_list(const FusedArgs &A, int tid, int n_threads) {
    for (int i = tid; i < A.n; i += n_threads) A.bl_xref[i] = A.xyzq[i];
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int n_pad = (A.n + 31) & ~31;
    for (int r = tid / 32; r < A.n; r += n_threads / 32) {
        const float4 xi = A.xyzq[r];
        int ex_lo = 0, ex_hi = 0;
        if (A.excl_start) {
            const int oi = A.orig[r];
            ex_lo = A.excl_start[oi];
            ex_hi = A.excl_start[oi + 1];
        }
        uint32_t *row = A.bl_list + (size_t)r * A.bl_stride;
        uint32_t cnt = 0;
        for (int j = lane; j < n_pad; j += 32) {
            bool hit = false;
            if (j < A.n && j != r) {
                const float4 xj = A.xyzq[j];
                float dx = __fsub_rn(xi.x, xj.x), dy = __fsub_rn(xi.y, xj.y), dz = __fsub_rn(xi.z, xj.z);
                if (PBC) {
                    dx = fused_min_image_exact(dx, A.p.ext[0], A.p.inv_ext[0]);
                    dy = fused_min_image_exact(dy, A.p.ext[1], A.p.inv_ext[1]);
                    dz = fused_min_image_exact(dz, A.p.ext[2], A.p.inv_ext[2]);
                }
                const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                hit = r2 < A.rl2;
                if (hit && ex_hi > ex_lo) {
                    const int oj = A.orig[j];
                    for (int e = ex_lo; e < ex_hi; ++e)
                        if (A.excl_idx[e] == oj) { hit = false; break; }
                }
            }
            const uint32_t mask = __ballot_sync(MC_FULL_MASK, hit);
            if (hit) row[cnt + (uint32_t)__popc(mask & lt)] = (uint32_t)j;
            cnt += (uint32_t)__popc(mask);
        }
        if (lane == 0) A.bl_count[r] = cnt;
    }
}

template <bool MULTI, int COUL, bool PBC>
__device__ __forceinline__ void fused_bonded(const FusedArgs &A, int tid, int n_threads) {
    const int n_bonded = A.bt.n_bonds + A.bt.n_angles + A.bt.n_dihedrals;
    for (int t = tid; t < n_bonded; t += n_threads) {
        float e, w;
        int kind;
        bonded_term_apply(t, A.bt, A.slot_of_orig, A.xyzq, A.p, A.force, e, w, kind);
    }
}

template <bool MULTI, int COUL, bool PBC>
__global__ void __launch_bounds__(FUSED_THREADS) md_fused_kernel(const FusedArgs A) {
    cg::grid_group grid = cg::this_grid();
    MC_DYN_SHARED_ALIGNED(unsigned char, s_raw, 16);
    // [brute: positions float4[n] | ] LJ table float2[T x T] [ | brute: types u16[n]]
    float4 *s_pos = reinterpret_cast<float4 *>(s_raw);
    float2 *s_tab = reinterpret_cast<float2 *>(s_raw + (A.brute ? sizeof(float4) * (size_t)A.n : 0));
    uint16_t *s_type = reinterpret_cast<uint16_t *>(s_tab + (MULTI ? A.p.n_types * A.p.n_types : 0));
    if (MULTI)
        for (int t = threadIdx.x; t < A.p.n_types * A.p.n_types; t += blockDim.x) s_tab[t] = A.ljtab[t];
    if (A.brute)
        for (int i = threadIdx.x; i < A.n; i += blockDim.x) s_type[i] = A.type[i];
    __syncthreads();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
    // every block keeps its own copy of the positions (brute mode), refreshed after each drift
    auto stage_positions = [&]() {
        __syncthreads();  // nobody still reads the old copy
        for (int i = threadIdx.x; i < A.n; i += blockDim.x) s_pos[i] = A.xyzq[i];
        __syncthreads();
    };
    const int n_bonded = A.bt.n_bonds + A.bt.n_angles + A.bt.n_dihedrals;
    int s = 0, flag = 0, rebuilds = 0;
    int n_dbg = 0;
#define MC_FUSED_STAMP() do { if (A.dbg && tid == 0 && n_dbg < 64) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); A.dbg[n_dbg++] = t_; } } while (0)
    MC_FUSED_STAMP();
    if (A.brute) {
        // (bl_keep: the private list of the previous launch still describes this system -- nothing but that launch has touched
        // positions, order or exclusions since -- and its displacement reference goes on counting)
        if (!A.bl_keep) {
            fused_build_list<PBC>(A, tid, n_threads);
            MC_FUSED_STAMP();
            grid.sync();
        }
        MC_FUSED_STAMP();
        if (A.need_forces) {
            stage_positions();
            fused_forces<MULTI, COUL, PBC>(A, s_tab, s_pos, s_type, tid, n_threads);
            if (n_bonded > 0) {
                grid.sync();
                fused_bonded<MULTI, COUL, PBC>(A, tid, n_threads);
            }
            grid.sync();
        }
    }
    for (; s < A.n_steps; ++s) {
        // ---- kick + drift (integrate.cu), displacement criterion without look-ahead
        const float kick = (s == 0 && A.first_half) ? 0.5f * A.dt : A.dt;
        // brute mode: the flag words of even and odd steps alternate, so the word of the next step can be cleared while
        // blocks may still be reading this one
        int *fw = A.brute ? A.bl_flags + (s & 1) : A.rebuild_flag;
        fused_kick_drift(A, kick, A.brute ? A.bl_xref : A.xref, fw, tid, n_threads);
        MC_FUSED_STAMP();
        grid.sync();
        MC_FUSED_STAMP();
        flag = *reinterpret_cast<volatile int *>(fw);
        if (A.brute) {
            if (tid == 0) A.bl_flags[(s + 1) & 1] = 0;
            if (flag & 2) break;  // grid-uniform: non-finite coordinates, the host reports
            if (flag & 1) {       // grid-uniform: an atom outran skin/2 -- new list from the positions just reached
                fused_build_list<PBC>(A, tid, n_threads);
                ++rebuilds;
                grid.sync();
            }
        } else if (flag & 3) {
            break;  // grid-uniform: this step ends after its drift; the host rebuilds / reports
        }
        if (A.brute) stage_positions();
        fused_forces<MULTI, COUL, PBC>(A, s_tab, s_pos, s_type, tid, n_threads);
        MC_FUSED_STAMP();
        if (n_bonded > 0) {
            grid.sync();  // every row is written before the bonded terms add to it
            fused_bonded<MULTI, COUL, PBC>(A, tid, n_threads);
        }
        grid.sync();
        MC_FUSED_STAMP();
    }
    if (s == A.n_steps) fused_half_kick(A, tid, n_threads);
    if (A.brute) {
        // is the HOST's list (reference positions A.xref) still good for the positions reached?  Only its next user cares.
        int stale = 0;
        if (A.xref) {
            for (int i = tid; i < A.n; i += n_threads) {
                const float4 x = A.xyzq[i], r = A.xref[i];
                const float dx = x.x - r.x, dy = x.y - r.y, dz = x.z - r.z;
                if (dx * dx + dy * dy + dz * dz > A.max_disp * A.max_disp) stale = 1;
            }
        }
        if (stale) atomicOr(A.bl_flags + 2, 1);
        grid.sync();
    }
    if (tid == 0) {
        A.out[0] = s == A.n_steps ? A.n_steps : s + 1;  // drifts completed; < n_steps or flagged: forces of the last one pending
        A.out[1] = A.brute ? (flag & 2) : flag;
        A.out[2] = A.brute ? *reinterpret_cast<volatile int *>(A.bl_flags + 2) : 0;
        A.out[3] = rebuilds;
    }
}

}  // namespace

// Largest system the fused kernel takes: the grid must be co-resident (cooperative launch) and the win is launch latency
int md_fused_max_atoms() { return 65536; }
// Largest system that keeps a private all-pairs list inside the kernel: rows of stride n (no row can overflow) = 4 n^2 bytes
int md_fused_brute_max_atoms() { return 2048; }

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
    const size_t smem = (MULTI ? sizeof(float2) * A.p.n_types * A.p.n_types : 0) +
                        (A.brute ? (sizeof(float4) + sizeof(uint16_t)) * (size_t)A.n + 16 : 0);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, md_fused_kernel<MULTI, COUL, PBC>, FUSED_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    // a grid barrier costs more the more blocks take part: no more blocks than the rows need, at most two per SM; a row is
    // spread over as many lanes (8, 16, 32) as still fit that grid -- at a thousand atoms 8 lanes per row would leave three
    // quarters of the SMs idle and two warps per scheduler to hide the latency of the gathers
    const int max_blocks = n_sms * std::min(per_sm, 2);
    int lanes = FUSED_LANES;
    while (lanes < 32 && (long long)A.n * (lanes * 2) <= (long long)max_blocks * FUSED_THREADS) lanes *= 2;
    if (A.lanes == 8 || A.lanes == 16 || A.lanes == 32) lanes = A.lanes;  // option fused_lanes
    const int want = (int)div_up((size_t)A.n * lanes, FUSED_THREADS);
    const int grid = std::max(1, std::min(want, max_blocks));
    FusedArgs a = A;
    a.lanes = lanes;
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
int md_fused_brute_max_atoms() { return 0; }
cudaError_t md_fused_prepare() { return cudaSuccess; }
cudaError_t launch_md_fused(const FusedArgs &, bool, int, bool, int, cudaStream_t, int64_t *) { return cudaErrorNotSupported; }
#endif
