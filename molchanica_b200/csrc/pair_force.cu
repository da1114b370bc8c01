// pair_force.cu -- the nonbonded pair-force inner loop: Lennard-Jones 12-6 + cutoff Coulomb over a
// full Verlet neighbour list.  Replaces lj_force_kernel / coulomb_force_kernel
// of reference src/cuda/cuda.cu:10-102 (all-pairs, dense per-pair sigma/eps arrays, RMW of out[i]
// in the inner loop) with:
//   - float4 xyzq records (one 16-byte gather per neighbour instead of float3 + charge)
//   - a T x T (sigma^2, 24 eps) table in shared memory instead of 8 B/pair of parameters in HBM
//   - LANES lanes per target atom reading LANES consecutive list entries (one 32-byte sector for
//     LANES = 8), partial forces reduced with warp shuffles -- no atomics, no RMW of the output
//   - full lists: every atom owns its row, so the result needs no scatter and no reverse halo
//
// Roofline: HBM.  Algorithmic bytes per launch = 32 N + 20 P_full: per atom 16 B xyzq
// read + 16 B (fx, fy, fz, e) write; per list entry 4 B index + 16 B gathered xyzq_j.
//
// Pair arithmetic follows reference src/cuda/util.cu:
//   LJ       sr = sigma/r; |F| = 24 eps (2 sr^12 - sr^6)/r; E = 4 eps (sr^12 - sr^6)      (:93-139)
//   Coulomb  F = dir * q_i q_j / (r^2 + 1e-6), dir = (r_i - r_j)/r                         (:54-63)
//   min image d -= rintf(d/ext)*ext                                                         (:65-71)
// The cutoff decision uses the oracle's exact fp32 r^2 (no fma) so that both sides mask the
// same pairs; everything after the mask is free to use fma / approximate reciprocals.
//
// Measured (profiles/): the kernel is instruction-issue bound, not HBM bound -- DRAM carries only
// the index stream, the 16-byte gathers hit L1/L2 -- so the variants trim instructions: the
// minimum image is applied only to rows of boundary-cell atoms (MC_FLAG_INTERIOR), the energy
// row sum only when a caller asks for it (ENERGY), and rows are cutoff-partitioned at build time.
#include "common.cuh"
#include "pair_force.cuh"
#include "pair_terms.cuh"

namespace {

// Defaults from the A/B of profiles/pair_force_r2t_knobs.txt (C4, one B200): four gathers in flight per lane and index
// loads that do not allocate in L1 (the list streams through once) -- 0.1536 ms against 0.1572 ms with two gathers and
// allocating loads; each knob alone is within the noise.  Accumulation order per lane is unchanged (same bits).
#ifndef MC_PF_DEPTH
#define MC_PF_DEPTH 4
#endif
#if !defined(MC_PF_IDX_ALLOC) && !defined(MC_PF_IDX_NOALLOC)
#define MC_PF_IDX_NOALLOC 1
#endif
__device__ __forceinline__ uint32_t pf_ld_idx(const uint32_t *p) {
#if defined(MC_PF_IDX_NOALLOC) && !defined(MC_HOST_SHIM)
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ float4 pf_ld_pos(const float4 *p) {
#if defined(MC_PF_POS_EVICT_LAST) && !defined(MC_HOST_SHIM)
    float4 v;
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// KSTR: distance between a lane's consecutive entries in units of LANES -- 1 for plain rows (entry k of a row at lst[k]),
// 4 for quad-interleaved rows (rows_interleave_*: the LANES-entry chunks of four consecutive rows alternate, so that the
// index load of a warp's four rows is ONE 128-byte line instead of four 32-byte sectors in four lines).
template <int LANES, bool MULTI, int COUL, bool WRAP, bool ENERGY, int KSTR>
__device__ __forceinline__ void row_loop(const float4 xi, const uint32_t *__restrict__ lst, uint32_t cnt, int sub,
                                         const float4 *__restrict__ xyzq, const uint16_t *__restrict__ type,
                                         const float2 *row, const NbParams &p, bool lj_on, Acc &a) {
    const float2 lj1 = make_float2(p.sig2, p.eps24);
    const float rc2_lj = lj_on ? p.rc2_lj : -1.f;
    uint32_t k = sub;
    lst += sub;
#if MC_PF_DEPTH != 2 || defined(MC_PF_IDX_NOALLOC) || defined(MC_PF_POS_EVICT_LAST)
    // A/B knobs (tools/gpu_r2_t.sh, profiles/pair_force_r2t_knobs.txt): MC_PF_DEPTH gathers in flight per lane, index loads that
    // do not allocate in L1 (the list streams through once), position loads marked evict-last
    for (; k + (MC_PF_DEPTH - 1) * LANES < cnt; k += MC_PF_DEPTH * LANES, lst += MC_PF_DEPTH * LANES * KSTR) {
        uint32_t j[MC_PF_DEPTH];
        float4 x[MC_PF_DEPTH];
        float2 l[MC_PF_DEPTH];
#pragma unroll
        for (int u = 0; u < MC_PF_DEPTH; ++u) j[u] = pf_ld_idx(lst + u * LANES * KSTR);
#pragma unroll
        for (int u = 0; u < MC_PF_DEPTH; ++u) x[u] = pf_ld_pos(xyzq + j[u]);
#pragma unroll
        for (int u = 0; u < MC_PF_DEPTH; ++u) l[u] = MULTI ? row[__ldg(type + j[u])] : lj1;
#pragma unroll
        for (int u = 0; u < MC_PF_DEPTH; ++u) pair_term<COUL, WRAP, ENERGY>(xi, x[u], l[u], p, rc2_lj, a);
    }
    for (; k < cnt; k += LANES, lst += LANES * KSTR) {
        const uint32_t j0 = pf_ld_idx(lst);
        const float4 x0 = pf_ld_pos(xyzq + j0);
        float2 l0 = lj1;
        if (MULTI) l0 = row[__ldg(type + j0)];
        pair_term<COUL, WRAP, ENERGY>(xi, x0, l0, p, rc2_lj, a);
    }
#else
    // two gathers in flight per lane
    for (; k + LANES < cnt; k += 2 * LANES, lst += 2 * LANES * KSTR) {
        const uint32_t j0 = __ldg(lst), j1 = __ldg(lst + LANES * KSTR);
        const float4 x0 = __ldg(xyzq + j0), x1 = __ldg(xyzq + j1);
        float2 l0 = lj1, l1 = lj1;
        if (MULTI) { l0 = row[__ldg(type + j0)]; l1 = row[__ldg(type + j1)]; }
        pair_term<COUL, WRAP, ENERGY>(xi, x0, l0, p, rc2_lj, a);
        pair_term<COUL, WRAP, ENERGY>(xi, x1, l1, p, rc2_lj, a);
    }
    if (k < cnt) {
        const uint32_t j0 = __ldg(lst);
        const float4 x0 = __ldg(xyzq + j0);
        float2 l0 = lj1;
        if (MULTI) l0 = row[__ldg(type + j0)];
        pair_term<COUL, WRAP, ENERGY>(xi, x0, l0, p, rc2_lj, a);
    }
#endif
}

// Warp-uniform variant of the row loop: all 32 lanes run the same trip count (the longest of the
// warp's 32/LANES rows), so a full-mask vote is legal inside it.  Rows are cutoff-partitioned at
// build time (inner entries first, skin shell last): once every lane of the warp is in its skin
// shell the vote fails and the whole warp skips the force arithmetic of that iteration.
template <int LANES, bool MULTI, int COUL, bool PBC, bool ENERGY>
__device__ __forceinline__ void row_loop_uniform(const float4 xi, const uint32_t *__restrict__ lst, uint32_t cnt, int sub,
                                                 bool wrap, const float4 *__restrict__ xyzq,
                                                 const uint16_t *__restrict__ type, const float2 *row, const NbParams &p,
                                                 bool lj_on, Acc &a) {
    uint32_t cnt_max = cnt;
#pragma unroll
    for (int d = LANES; d < 32; d <<= 1) cnt_max = max(cnt_max, __shfl_xor_sync(MC_FULL_MASK, cnt_max, d));
    const bool any_wrap = PBC && __any_sync(MC_FULL_MASK, wrap);
    const float2 lj1 = make_float2(p.sig2, p.eps24);
    const float rc2_max = fmaxf(p.rc2_lj, COUL != MC_COULOMB_NONE ? p.rc2_q : 0.f);
    for (uint32_t kk = 0; kk < cnt_max; kk += LANES) {
        const uint32_t k = kk + sub;
        const bool valid = k < cnt;
        float dx = 0.f, dy = 0.f, dz = 0.f, r2 = 3.0e38f;
        float4 xj = xi;
        float2 lj = lj1;
        if (valid) {
            const uint32_t j = __ldg(lst + k);
            xj = __ldg(xyzq + j);
            if (MULTI) lj = row[__ldg(type + j)];
            dx = xi.x - xj.x; dy = xi.y - xj.y; dz = xi.z - xj.z;
            if (any_wrap && wrap) {
                dx = __fmaf_rn(-rintf(dx * p.inv_ext[0]), p.ext[0], dx);
                dy = __fmaf_rn(-rintf(dy * p.inv_ext[1]), p.ext[1], dy);
                dz = __fmaf_rn(-rintf(dz * p.inv_ext[2]), p.ext[2], dz);
            }
            r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        }
        if (!__any_sync(MC_FULL_MASK, r2 < rc2_max)) continue;  // the whole warp is in its skin shell
        float f = 0.f, e = 0.f;
        if (lj_on && r2 < p.rc2_lj) {
            const float ir2 = rcp_approx(r2);
            const float s2 = lj.x * ir2;
            const float s6 = s2 * s2 * s2;
            f = lj.y * s6 * __fmaf_rn(2.f, s6, -1.f) * ir2;
            if (ENERGY) e = lj.y * (1.f / 6.f) * s6 * (s6 - 1.f);
        }
        if (COUL != MC_COULOMB_NONE) {
            if (r2 < p.rc2_q) {
                const float qq = xi.w * xj.w;
                const float ir = rsqrt_approx(r2);
                if (COUL == MC_COULOMB_PLAIN) {
                    f = __fmaf_rn(qq * ir, rcp_approx(r2 + MC_SOFTENING_SQ), f);
                    if (ENERGY) e = __fmaf_rn(qq, ir, e);
                } else {
                    const float r = r2 * ir;
                    const float ar = p.alpha * r;
                    const float erfc_ar = erfcf(ar);
                    const float ex = __expf(-ar * ar);
                    const float ir2 = ir * ir;
                    f = __fmaf_rn(qq * ir, __fmaf_rn(erfc_ar, ir2, 2.f * p.alpha * MC_INV_SQRT_PI * ex * ir), f);
                    if (ENERGY) e = __fmaf_rn(qq * erfc_ar, ir, e);
                }
            }
        }
        a.fx = __fmaf_rn(dx, f, a.fx);
        a.fy = __fmaf_rn(dy, f, a.fy);
        a.fz = __fmaf_rn(dz, f, a.fz);
        if (ENERGY) a.e += e;
    }
}

// ILV: nbr_start / nbr_list are the quad-interleaved copy of the rows (base of every quad of four slots, see
// rows_interleave_sizes_kernel); nbr_count is the rows' own.
template <int LANES, bool MULTI, int COUL, bool PBC, bool ENERGY, bool UNIFORM, bool ILV = false>
__global__ void __launch_bounds__(128) pair_force_kernel(int n_rows, int row0, const float4 *__restrict__ xyzq,
                                                          const uint16_t *__restrict__ type,
                                                          const uint8_t *__restrict__ flags,
                                                          const uint32_t *__restrict__ nbr_start,
                                                          const uint32_t *__restrict__ nbr_count,
                                                          const uint32_t *__restrict__ nbr_list,
                                                          const float2 *__restrict__ ljtab, const NbParams p,
                                                          const int lj_on, float4 *__restrict__ force,
                                                          const int n_interior, const int n_first, const HaloWait hw) {
    MC_DYN_SHARED(float2, s_tab);
    constexpr int ROWS_PER_BLOCK = 128 / LANES;
    if (hw.ready_prev && (int)(blockIdx.x + 1) * ROWS_PER_BLOCK > n_interior) {
        // Decomposed rank, fused halo (halo_sync.cuh): launch rows are ordered interior first, then the first and
        // the last owned layer, whose rows gather ghosts that the neighbours' kick_drift kernels store over
        // NVLink.  Blocks of those rows wait for this epoch's ready flags -- by then the interior rows have
        // covered the latency -- and then drop this SM's L1 (a gpu-scope fence invalidates it): an interior
        // block may have pulled in a 32-byte sector that a ghost shares with an owned atom before the push landed.
        if (threadIdx.x == 0) {
            halo_spin(hw.ready_prev, hw.want, hw.err);
            halo_spin(hw.ready_next, hw.want, hw.err);
            __threadfence();
        }
        __syncthreads();
    }
    if (MULTI) {
        for (int t = threadIdx.x; t < p.n_types * p.n_types; t += blockDim.x) s_tab[t] = ljtab[t];
        __syncthreads();
    }
    const int sub = threadIdx.x % LANES;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const bool live = r < n_rows;
    // r < n_interior: interior row n_first + r; then the first layer's rows 0 .. n_first-1; then the last layer's
    const int i = row0 + (r < n_interior ? r + n_first : (r < n_interior + n_first ? r - n_interior : r));
    Acc a = {0.f, 0.f, 0.f, 0.f};
    if (UNIFORM) {
        // every lane of the warp enters the loop (dead rows with a zero count)
        const int ii = live ? i : row0;
        const float4 xi = __ldg(xyzq + ii);
        const uint32_t start = __ldg(nbr_start + ii), cnt = live ? __ldg(nbr_count + ii) : 0u;
        const float2 *row = MULTI ? s_tab + (int)__ldg(type + ii) * p.n_types : nullptr;
        const bool wrap = PBC && !(__ldg(flags + ii) & MC_FLAG_INTERIOR);
        row_loop_uniform<LANES, MULTI, COUL, PBC, ENERGY>(xi, nbr_list + start, cnt, sub, wrap, xyzq, type, row, p, lj_on, a);
    } else if (live) {
        const float4 xi = __ldg(xyzq + i);
        const uint32_t start = ILV ? __ldg(nbr_start + (i >> 2)) + (uint32_t)(i & 3) * LANES : __ldg(nbr_start + i);
        const uint32_t cnt = __ldg(nbr_count + i);
        const float2 *row = MULTI ? s_tab + (int)__ldg(type + i) * p.n_types : nullptr;
        const uint32_t *lst = nbr_list + start;
        constexpr int KSTR = ILV ? 4 : 1;
        // Atoms of interior cells never see a wrapped neighbour between two list builds: their raw
        // differences already are the minimum image (n == 0), so the 9-instruction wrap is skipped.
        const bool wrap = PBC && !(__ldg(flags + i) & MC_FLAG_INTERIOR);
        if (wrap) row_loop<LANES, MULTI, COUL, true, ENERGY, KSTR>(xi, lst, cnt, sub, xyzq, type, row, p, lj_on, a);
        else row_loop<LANES, MULTI, COUL, false, ENERGY, KSTR>(xi, lst, cnt, sub, xyzq, type, row, p, lj_on, a);
    }
    // warp-shuffle partial-force reduction across the LANES lanes of this row
#pragma unroll
    for (int d = LANES / 2; d > 0; d >>= 1) {
        a.fx += __shfl_xor_sync(MC_FULL_MASK, a.fx, d);
        a.fy += __shfl_xor_sync(MC_FULL_MASK, a.fy, d);
        a.fz += __shfl_xor_sync(MC_FULL_MASK, a.fz, d);
        if (ENERGY) a.e += __shfl_xor_sync(MC_FULL_MASK, a.e, d);
    }
    if (live && sub == 0) force[i] = make_float4(a.fx, a.fy, a.fz, a.e);
}

// Amber 1-4 rows: every atom sums its own 1-4 partners (deterministic, no atomics), no cutoff,
// LJ x scale_lj, Coulomb x scale_q.  Partner ids are original ids.
__global__ void __launch_bounds__(128) pairs14_kernel(int n_rows, int row0, const float4 *__restrict__ xyzq,
                                                       const uint16_t *__restrict__ type, const int *__restrict__ orig,
                                                       const int *__restrict__ slot_of_orig,
                                                       const int32_t *__restrict__ p14_start,
                                                       const int32_t *__restrict__ p14_idx,
                                                       const float2 *__restrict__ ljtab, const NbParams p,
                                                       float scale_lj, float scale_q, int lj_on, int coul_on,
                                                       float4 *__restrict__ force) {
    const int kr = blockIdx.x * blockDim.x + threadIdx.x;
    if (kr >= n_rows) return;
    const int k = row0 + kr;
    const int oi = orig[k];
    const int e0 = p14_start[oi], e1 = p14_start[oi + 1];
    if (e0 == e1) return;
    const float4 xi = xyzq[k];
    const int ti = type[k];
    float4 acc = force[k];
    for (int e = e0; e < e1; ++e) {
        const int j = slot_of_orig[p14_idx[e]];
        if (j < 0) continue;  // partner not held by this rank (cannot happen within one ghost layer)
        const float4 xj = xyzq[j];
        float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
        if (p.periodic) {
            dx -= rintf(dx * p.inv_ext[0]) * p.ext[0];
            dy -= rintf(dy * p.inv_ext[1]) * p.ext[1];
            dz -= rintf(dz * p.inv_ext[2]) * p.ext[2];
        }
        const float r2 = dx * dx + dy * dy + dz * dz;
        const float ir2 = 1.0f / r2;
        float f = 0.f, en = 0.f;
        if (lj_on) {
            const float2 lj = ljtab[ti * p.n_types + type[j]];
            const float s2 = lj.x * ir2, s6 = s2 * s2 * s2;
            f += scale_lj * lj.y * s6 * (2.f * s6 - 1.f) * ir2;
            en += scale_lj * lj.y * (1.f / 6.f) * s6 * (s6 - 1.f);
        }
        if (coul_on) {
            const float qq = xi.w * xj.w, ir = rsqrtf(r2);
            f += scale_q * qq * ir / (r2 + MC_SOFTENING_SQ);
            en += scale_q * qq * ir;
        }
        acc.x += dx * f; acc.y += dy * f; acc.z += dz * f; acc.w += en;
    }
    force[k] = acc;
}

// Quad-interleaved copy of the rows for pair_force_kernel<LANES = 8, ..., ILV> (once per list build).  The slots are taken
// four at a time; a quad's rows are cut into chunks of 8 entries and stored chunk-major: [row0 c0 | row1 c0 | row2 c0 |
// row3 c0 | row0 c1 | ...], every row padded to the quad's longest one (entries past a row's count are never read).  A
// warp of the force kernel -- four rows x 8 lanes -- then loads its 32 indices from ONE 128-byte line; with plain rows the
// same load touches four lines, and the kernel is bound by exactly those L1 wavefronts.
// Pass 1: chunks per quad, one block-wide scan, one claim per block on *cursor (units: entries); qbase[q] = first entry.
constexpr int ILV_CHUNK = 8;
__global__ void __launch_bounds__(256) rows_interleave_sizes_kernel(int n_slots, const uint32_t *__restrict__ nbr_count,
                                                                     uint32_t *__restrict__ qbase, uint32_t *__restrict__ cursor) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    const int n_quads = (n_slots + 3) >> 2;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t mx = 0;
    if (q < n_quads) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = 4 * q + r;
            if (i < n_slots) mx = max(mx, nbr_count[i]);
        }
    }
    const uint32_t mine = ((mx + ILV_CHUNK - 1) / ILV_CHUNK) * (4 * ILV_CHUNK);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(MC_FULL_MASK, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int k = 0; k < 8; ++k) { const uint32_t t = s_warp[k]; s_warp[k] = tot; tot += t; }
        s_base = tot ? atomicAdd(cursor, tot) : 0u;
    }
    __syncthreads();
    if (q < n_quads) qbase[q] = s_base + s_warp[w] + incl - mine;
}

// Pass 2: one warp per quad copies the four rows chunk by chunk (reads: four 32-byte sectors, writes: one 128-byte line).
__global__ void __launch_bounds__(256) rows_interleave_copy_kernel(int n_slots, const uint32_t *__restrict__ nbr_start,
                                                                    const uint32_t *__restrict__ nbr_count,
                                                                    const uint32_t *__restrict__ nbr_list,
                                                                    const uint32_t *__restrict__ qbase, uint32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int q = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int n_quads = (n_slots + 3) >> 2;
    if (q >= n_quads) return;  // whole warps
    const int i = 4 * q + (lane >> 3), sub = lane & 7;
    uint32_t cnt = 0, start = 0;
    if (i < n_slots) { cnt = nbr_count[i]; start = nbr_start[i]; }
    uint32_t mx = cnt;
    mx = max(mx, __shfl_xor_sync(MC_FULL_MASK, mx, 8));
    mx = max(mx, __shfl_xor_sync(MC_FULL_MASK, mx, 16));
    uint32_t *dst = out + qbase[q] + lane;
    for (uint32_t k = sub; k - sub < mx; k += ILV_CHUNK, dst += 4 * ILV_CHUNK)
        *dst = k < cnt ? __ldg(nbr_list + start + k) : 0u;
}

// Two-stage deterministic reduction: per-block partials {sum e_i, sum 1/2 m v^2, n_mobile}.
constexpr int RED_BLOCKS = 592;  // 4 x 148 SMs
__global__ void __launch_bounds__(256) energy_partial_kernel(int n_rows, const float4 *__restrict__ force,
                                                              const float4 *__restrict__ vel, const uint8_t *__restrict__ flags,
                                                              double *__restrict__ partial) {
    double e = 0.0, ke = 0.0, nm = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += gridDim.x * blockDim.x) {
        e += (double)force[i].w;
        const float4 v = vel[i];
        if (v.w > 0.f && !(flags && (flags[i] & MC_FLAG_STATIC))) {  // a static atom is no degree of freedom, whatever its mass
            ke += 0.5 * ((double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z) / (double)v.w;
            nm += 1.0;
        }
    }
    __shared__ double s[3][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        e += __shfl_xor_sync(MC_FULL_MASK, e, d);
        ke += __shfl_xor_sync(MC_FULL_MASK, ke, d);
        nm += __shfl_xor_sync(MC_FULL_MASK, nm, d);
    }
    if (lane == 0) { s[0][w] = e; s[1][w] = ke; s[2][w] = nm; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += s[threadIdx.x][k];
        partial[threadIdx.x * RED_BLOCKS + blockIdx.x] = t;
    }
}

__global__ void energy_final_kernel(const double *__restrict__ partial, int nb, double *__restrict__ out) {
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int k = 0; k < nb; ++k) t += partial[threadIdx.x * RED_BLOCKS + k];
        out[threadIdx.x] = t;
    }
}

#ifdef MC_HAVE_LAUNCH  // the stand-ins of tests/cpp/shim/ and shim_mt/ have no launcher; shim_fiber/ has
template <int LANES>
void launch_lanes(const PairLaunch &L, cudaStream_t st) {
    const int rows_per_block = 128 / LANES;
    const unsigned blocks = div_up(L.n_rows, rows_per_block);
    const size_t smem = L.multi ? sizeof(float2) * L.p.n_types * L.p.n_types : 0;
#define MC_PF(M, C, P, E)                                                                                               \
    do {                                                                                                                \
        if (LANES == 8 && L.ilv_list && !L.uniform)                                                                     \
            MC_LAUNCH(pair_force_kernel<8 MC_COMMA M MC_COMMA C MC_COMMA P MC_COMMA E MC_COMMA false MC_COMMA true>, blocks, 128, smem, st, \
                      L.n_rows, L.row0, L.xyzq, L.type, L.flags, L.ilv_qbase, L.nbr_count, L.ilv_list, L.ljtab, L.p, L.lj_on,   \
                      L.force, L.n_interior, L.n_first, L.wait);                                                        \
        else if (L.uniform)                                                                                                  \
            MC_LAUNCH(pair_force_kernel<LANES MC_COMMA M MC_COMMA C MC_COMMA P MC_COMMA E MC_COMMA true>, blocks, 128, smem, st, \
                      L.n_rows, L.row0, L.xyzq, L.type, L.flags, L.nbr_start, L.nbr_count, L.nbr_list, L.ljtab, L.p, L.lj_on,   \
                      L.force, L.n_interior, L.n_first, L.wait);                                                        \
        else                                                                                                            \
            MC_LAUNCH(pair_force_kernel<LANES MC_COMMA M MC_COMMA C MC_COMMA P MC_COMMA E MC_COMMA false>, blocks, 128, smem, st, \
                      L.n_rows, L.row0, L.xyzq, L.type, L.flags, L.nbr_start, L.nbr_count, L.nbr_list, L.ljtab, L.p, L.lj_on,   \
                      L.force, L.n_interior, L.n_first, L.wait);                                                        \
    } while (0)
#define MC_PF_E(M, C, P) \
    if (L.energy) MC_PF(M, C, P, true); else MC_PF(M, C, P, false)
#define MC_PF_C(M, P)                                                    \
    switch (L.coul) {                                                    \
        case MC_COULOMB_NONE: MC_PF_E(M, MC_COULOMB_NONE, P); break;     \
        case MC_COULOMB_PLAIN: MC_PF_E(M, MC_COULOMB_PLAIN, P); break;   \
        default: MC_PF_E(M, MC_COULOMB_ERFC, P); break;                  \
    }
    if (L.multi) { if (L.p.periodic) { MC_PF_C(true, true) } else { MC_PF_C(true, false) } }
    else { if (L.p.periodic) { MC_PF_C(false, true) } else { MC_PF_C(false, false) } }
#undef MC_PF_C
#undef MC_PF_E
#undef MC_PF
}

#endif  // MC_HOST_SHIM

}  // namespace

#ifdef MC_HAVE_LAUNCH  // the stand-ins of tests/cpp/shim/ and shim_mt/ have no launcher; shim_fiber/ has
int pair_force_max_types() { return 160; }  // 160^2 * 8 B = 200 KB of the 227 KB shared memory

cudaError_t pair_force_prepare() {
    // opt in to large dynamic shared memory for the multi-type instantiations
    cudaError_t e = cudaSuccess;
#define MC_ATTR(LN, C, P, E)                                                                                     \
    if (e == cudaSuccess)                                                                                        \
        e = cudaFuncSetAttribute(pair_force_kernel<LN, true, C, P, E, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 200 * 1024);                                                                    \
    if (e == cudaSuccess)                                                                                        \
        e = cudaFuncSetAttribute(pair_force_kernel<LN, true, C, P, E, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 200 * 1024);                                                                    \
    if (e == cudaSuccess && LN == 8)                                                                             \
        e = cudaFuncSetAttribute(pair_force_kernel<8, true, C, P, E, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 200 * 1024);
#define MC_ATTR_C(LN, C) MC_ATTR(LN, C, true, true) MC_ATTR(LN, C, true, false) MC_ATTR(LN, C, false, true) MC_ATTR(LN, C, false, false)
#define MC_ATTR_L(LN) MC_ATTR_C(LN, MC_COULOMB_NONE) MC_ATTR_C(LN, MC_COULOMB_PLAIN) MC_ATTR_C(LN, MC_COULOMB_ERFC)
    MC_ATTR_L(4) MC_ATTR_L(8) MC_ATTR_L(16) MC_ATTR_L(32)
#undef MC_ATTR_L
#undef MC_ATTR_C
#undef MC_ATTR
    return e;
}

void launch_pair_force(const PairLaunch &L, cudaStream_t st, int64_t *launches) {
    if (L.n_rows <= 0) return;
    switch (L.lanes) {
        case 4: launch_lanes<4>(L, st); break;
        case 16: launch_lanes<16>(L, st); break;
        case 32: launch_lanes<32>(L, st); break;
        default: launch_lanes<8>(L, st); break;
    }
    *launches += 1;
}

void launch_pairs14(int n_rows, int row0, const float4 *xyzq, const uint16_t *type, const int *orig, const int *slot_of_orig,
                    const int32_t *p14_start, const int32_t *p14_idx, const float2 *ljtab, const NbParams &p,
                    float scale_lj, float scale_q, int lj_on, int coul_on, float4 *force, cudaStream_t st,
                    int64_t *launches) {
    if (n_rows <= 0) return;
    MC_LAUNCH(pairs14_kernel, div_up(n_rows, 128), 128, 0, st, n_rows, row0, xyzq, type, orig, slot_of_orig, p14_start, p14_idx, ljtab,
                                                       p, scale_lj, scale_q, lj_on, coul_on, force);
    *launches += 1;
}

void launch_rows_interleave_sizes(int n_slots, const uint32_t *nbr_count, uint32_t *qbase, uint32_t *cursor, cudaStream_t st, int64_t *launches) {
    if (n_slots <= 0) return;
    MC_LAUNCH(rows_interleave_sizes_kernel, div_up((size_t)((n_slots + 3) >> 2), 256), 256, 0, st, n_slots, nbr_count, qbase, cursor);
    *launches += 1;
}

void launch_rows_interleave_copy(int n_slots, const uint32_t *nbr_start, const uint32_t *nbr_count, const uint32_t *nbr_list,
                                 const uint32_t *qbase, uint32_t *out, cudaStream_t st, int64_t *launches) {
    if (n_slots <= 0) return;
    MC_LAUNCH(rows_interleave_copy_kernel, div_up((size_t)((n_slots + 3) >> 2) * 32, 256), 256, 0, st, n_slots, nbr_start, nbr_count, nbr_list,
              qbase, out);
    *launches += 1;
}

int energy_partial_elems() { return 3 * RED_BLOCKS; }

void launch_energy_reduce(int n_rows, const float4 *force, const float4 *vel, const uint8_t *flags, double *partial, double *out3,
                          cudaStream_t st, int64_t *launches) {
    int nb = (int)div_up(n_rows > 0 ? n_rows : 1, 256);
    if (nb > RED_BLOCKS) nb = RED_BLOCKS;
    MC_LAUNCH(energy_partial_kernel, nb, 256, 0, st, n_rows, force, vel, flags, partial);
    MC_LAUNCH(energy_final_kernel, 1, 32, 0, st, partial, nb, out3);
    *launches += 2;
}
#endif  // MC_HOST_SHIM
