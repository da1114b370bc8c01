// tile_build.cu -- single-pass Verlet-list build with the candidate tile staged in shared memory
// by TMA (the production path; neighbor.cu's two-pass global sweep is the fallback for
// neighbourhoods that do not fit shared memory).
//
// Persistent, warp-specialised kernel: as many CTAs as fit the chip, each pulling (cell, slice)
// work items from a global counter.  The 27-cell neighbourhood of a cell is at most 9 (+ periodic
// wrap splits) CONTIGUOUS slot ranges of the cell-ordered xyzq array, so the whole candidate tile
// is brought on chip by 1-D bulk-tensor copies (cp.async.bulk.shared::cluster.global, SASS UBLKCP)
// that complete on an mbarrier -- no register staging, no per-thread loads:
//   warp 0 (producer)    : claims the NEXT work item, looks up its ranges, waits for a free stage
//                          (empty barrier), arms the full barrier with the byte count and issues the
//                          bulk copies; also writes the slot ids of the staged atoms
//   warps 1..8 (consumers): wait on the full barrier, sweep the staged tile out of shared memory for
//                          1..4 atoms each (conflict-free LDS.128, lanes = consecutive candidates,
//                          one tile load shared by all atoms of the warp), then release the stage
// so the copy of item k+1 overlaps the sweep of item k (2-stage ring).
//
// Single pass: a warp records the accept decisions of its atoms as per-lane bit masks, reduces
// them to row lengths, claims room for its rows with ONE atomicAdd on the list cursor (rows padded
// to 8 entries = 32-byte sectors) and writes the rows.  Row placement in nbr_list therefore follows
// completion order, but every row's CONTENT and ORDER are deterministic (chunk-major, lane-major),
// so forces are bit-reproducible; mc_get_neighbors exports rows sorted by atom id regardless.
// Each row is partitioned: entries already inside the force cutoff at build time at the front,
// skin-shell entries at the back, so the tail iterations of the force kernel fail the cutoff test
// warp-wide.  The accept test is the oracle's fp32 expression bit for bit (see neighbor.cu).
#include <algorithm>

#include "common.cuh"
#include "neighbor.cuh"
#include "pair_terms.cuh"  // packed fp32 helpers
#include "tile_ring.cuh"

namespace {

constexpr int TILE_WARPS = 8;   // consumer warps
constexpr int TILE_STAGES = 2;  // tiles in flight per CTA
constexpr int TILE_A = 4;       // atoms swept together by one warp

#ifndef MC_HOST_SHIM
__device__ __forceinline__ uint32_t mc_brev(uint32_t v) { return __brev(v); }
#else
inline uint32_t mc_brev(uint32_t v) {
    uint32_t r = 0;
    for (int b = 0; b < 32; ++b) r |= ((v >> b) & 1u) << (31 - b);
    return r;
}
#endif

__device__ __forceinline__ float min_image_exact(float d, float ext, float inv_ext) {
    float q = __fmul_rn(d, inv_ext);
    float n = rintf(q);
    if (fabsf(q - n) > 0.4999f) n = rintf(__fdiv_rn(d, ext));
    return __fmaf_rn(-n, ext, d);
}

#ifndef MC_HOST_SHIM
// barrier 1: the consumer warps only (the producer never joins it)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TILE_WARPS * 32) : "memory"); }
#else
inline void consumer_sync() { shim_named_barrier(1, TILE_WARPS * 32); }
#endif

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t *total) {
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(MC_FULL_MASK, inc, d);
        if (lane >= d) inc += t;
    }
    *total = __shfl_sync(MC_FULL_MASK, inc, 31);
    return inc - v;
}

struct StageMeta {
    uint32_t m;         // staged atoms (the tile is padded with NaN positions to whole chunks)
    uint32_t a0, a1;    // atoms of the work item's cell (0xffffffff = no more work)
    uint32_t slice;     // which slice of the cell's atoms this item covers
    uint32_t self_off;  // tile index of atom a0 (the own cell is part of the tile)
    int wrap;           // 0: the cell's stencil never wraps; 1: wraps, >= 3 cells per axis; 2: wraps, tiny grid (exact path)
};

// Candidates are swept in chunks of 32 x B (B odd, <= 31): lane L owns the B CONSECUTIVE tile entries
// t0 + L*B .. t0 + L*B + B-1, so "lane-major" order is plain tile order (rows keep the memory
// locality of the cell-sorted atoms), and the odd stride keeps the 16-byte shared-memory reads of a
// quarter warp on distinct banks.
__device__ __forceinline__ uint32_t chunk_B(uint32_t remaining) {
    if (remaining >= 992u) return 31u;
    uint32_t b = (remaining + 31u) >> 5;
    if (b == 0u) b = 1u;
    return b | 1u;
}
__device__ __forceinline__ uint32_t padded_size(uint32_t m) {  // tile entries incl. NaN padding
    const uint32_t full = m / 992u, rem = m - full * 992u;
    return full * 992u + (rem ? 32u * chunk_B(rem) : 0u);
}

// Accept decisions of NA atoms (NA = 2 or 4: always whole PAIRS of atoms, a missing one is parked on NaN) against one chunk:
// bit `it` of hit[k] / inn[k] (per lane) = candidate t0 + lane*B + it is listed / is inside the force cutoff.  Two atoms
// share every instruction of the distance (packed fp32: FADD2 / FMUL2 with the candidate broadcast; the two additions of
// r^2 stay scalar so that ptxas cannot contract them with the products -- the expression must round exactly like the
// oracle's ((dx*dx)+(dy*dy))+(dz*dz)).  The decisions are shifted into the masks (2 instructions per test instead of a
// variable shift + select + or) and bit-reversed once per chunk; the own atom's bit is cleared afterwards instead of being
// tested per candidate.  WRAP is warp-uniform: cells whose stencil does not wrap skip the minimum image -- for such a cell
// every candidate's raw difference either IS the minimum image (|d| <= ext/2, n == 0) or belongs to a pair whose nearest
// image is beyond the list radius as well, so the decision is unchanged.  Parked atoms and tile padding are NaN: never
// accepted.  PART: also record which hits are inside the force cutoff (rows partitioned inner / skin shell for the
// warp-uniform pair loop); off by default.
template <int NA, int WRAP, bool PART>
__device__ __forceinline__ void sweep_chunk(const float4 *tile, uint32_t t0, uint32_t B, const GridParams &g, float rl2,
                                            float rc2_inner, const float4 (&pi)[TILE_A], const uint32_t (&t_self)[TILE_A],
                                            int lane, uint32_t (&hit)[TILE_A], uint32_t (&inn)[TILE_A]) {
#pragma unroll
    for (int k = 0; k < NA; ++k) hit[k] = inn[k] = 0u;
    const uint32_t tl = t0 + (uint32_t)lane * B;
    float2 px[NA / 2], py[NA / 2], pz[NA / 2];
#pragma unroll
    for (int h = 0; h < NA / 2; ++h) {
        px[h] = make_float2(pi[2 * h].x, pi[2 * h + 1].x);
        py[h] = make_float2(pi[2 * h].y, pi[2 * h + 1].y);
        pz[h] = make_float2(pi[2 * h].z, pi[2 * h + 1].z);
    }
    for (uint32_t it = 0; it < B; ++it) {
        const float4 pj = tile[tl + it];
        const float2 nx = make_float2(-pj.x, -pj.x), ny = make_float2(-pj.y, -pj.y), nz = make_float2(-pj.z, -pj.z);
#pragma unroll
        for (int h = 0; h < NA / 2; ++h) {
            float2 dx = mc_add2(px[h], nx), dy = mc_add2(py[h], ny), dz = mc_add2(pz[h], nz);
            if (WRAP == 2) {
                dx.x = min_image_exact(dx.x, g.ext[0], g.inv_ext[0]); dx.y = min_image_exact(dx.y, g.ext[0], g.inv_ext[0]);
                dy.x = min_image_exact(dy.x, g.ext[1], g.inv_ext[1]); dy.y = min_image_exact(dy.y, g.ext[1], g.inv_ext[1]);
                dz.x = min_image_exact(dz.x, g.ext[2], g.inv_ext[2]); dz.y = min_image_exact(dz.y, g.ext[2], g.inv_ext[2]);
            } else if (WRAP == 1) {
                // >= 3 cells per axis: ext >= 3 r_list, so whenever rintf(d * inv_ext) could differ from
                // rintf(d / ext) (|d| within rounding of ext/2) both images lie beyond 1.4 r_list and the
                // decision is the same; everywhere else the two agree and d - n*ext is exact
                const float2 qx = mc_mul2(dx, make_float2(g.inv_ext[0], g.inv_ext[0])), qy = mc_mul2(dy, make_float2(g.inv_ext[1], g.inv_ext[1])),
                             qz = mc_mul2(dz, make_float2(g.inv_ext[2], g.inv_ext[2]));
                dx = mc_fma2(make_float2(-rintf(qx.x), -rintf(qx.y)), make_float2(g.ext[0], g.ext[0]), dx);
                dy = mc_fma2(make_float2(-rintf(qy.x), -rintf(qy.y)), make_float2(g.ext[1], g.ext[1]), dy);
                dz = mc_fma2(make_float2(-rintf(qz.x), -rintf(qz.y)), make_float2(g.ext[2], g.ext[2]), dz);
            }
            const float2 sx = mc_mul2(dx, dx), sy = mc_mul2(dy, dy), sz = mc_mul2(dz, dz);
            const float r2a = __fadd_rn(__fadd_rn(sx.x, sy.x), sz.x), r2b = __fadd_rn(__fadd_rn(sx.y, sy.y), sz.y);
            hit[2 * h] = (hit[2 * h] << 1) | (r2a < rl2 ? 1u : 0u);
            hit[2 * h + 1] = (hit[2 * h + 1] << 1) | (r2b < rl2 ? 1u : 0u);
            if (PART) {
                inn[2 * h] = (inn[2 * h] << 1) | (r2a < rc2_inner ? 1u : 0u);
                inn[2 * h + 1] = (inn[2 * h + 1] << 1) | (r2b < rc2_inner ? 1u : 0u);
            }
        }
    }
    // candidate `it` sits in bit B-1-it: reverse, then clear the bit of the atom itself
    const uint32_t sh = 32u - B;
#pragma unroll
    for (int k = 0; k < NA; ++k) {
        hit[k] = mc_brev(hit[k]) >> sh;
        if (PART) inn[k] = mc_brev(inn[k]) >> sh;
        const uint32_t ds = t_self[k] - tl;  // (huge when the own atom is not in this lane's entries)
        if (ds < B) hit[k] &= ~(1u << ds);
    }
}

// Drop excluded partners (1-2 / 1-3 / 1-4, original ids) from the hit masks.  Rare: only atoms that
// carry exclusions pay for it, and only for their hits.
template <int NA>
__device__ __forceinline__ void apply_exclusions(const uint32_t *tile_slot, uint32_t tl /* t0 + lane*B */,
                                                 const int (&ex_lo)[TILE_A], const int (&ex_hi)[TILE_A],
                                                 const int *__restrict__ orig, const int32_t *__restrict__ excl_idx,
                                                 uint32_t (&hit)[TILE_A]) {
#pragma unroll
    for (int k = 0; k < NA; ++k) {
        if (ex_hi[k] <= ex_lo[k]) continue;
        uint32_t mleft = hit[k];
        while (mleft) {
            const int it = __ffs(mleft) - 1;
            mleft &= mleft - 1u;
            const int oj = orig[tile_slot[tl + (uint32_t)it]];
            for (int e = ex_lo[k]; e < ex_hi[k]; ++e)
                if (excl_idx[e] == oj) { hit[k] &= ~(1u << it); break; }
        }
    }
}

struct RowState {
    float4 pi[TILE_A];
    uint32_t t_self[TILE_A];
    int ex_lo[TILE_A], ex_hi[TILE_A];
    uint32_t hit[TILE_A], inn[TILE_A];          // masks of the (single / current) chunk
    uint32_t lane_in[TILE_A], lane_out[TILE_A];  // exclusive lane prefixes of the per-lane totals
    uint32_t len[TILE_A], n_inner[TILE_A];
};

// phase 1: accept decisions and row lengths of the NA atoms of this warp
template <int NA, int WRAP, bool PART>
__device__ __forceinline__ void rows_phase1(const float4 *tile, const uint32_t *tile_slot, const StageMeta &M, const GridParams &g,
                                            float rl2, float rc2_inner, const uint32_t (&ia)[TILE_A],
                                            const float4 *__restrict__ xyzq, const int *__restrict__ orig,
                                            const int32_t *__restrict__ excl_start, const int32_t *__restrict__ excl_idx,
                                            int lane, int na, RowState &R) {
    const float qnan = __int_as_float(0x7fffffff);
#pragma unroll
    for (int k = 0; k < TILE_A; ++k) {
        R.pi[k] = make_float4(qnan, qnan, qnan, 0.f);
        R.t_self[k] = 0xffffffffu;
        R.ex_lo[k] = R.ex_hi[k] = 0;
        R.hit[k] = R.inn[k] = 0u;
        R.len[k] = R.n_inner[k] = 0u;
        if (k < NA && k < na) {
            R.t_self[k] = M.self_off + (ia[k] - M.a0);
            R.pi[k] = tile[R.t_self[k]];  // the own cell is part of the staged tile
            if (excl_start) {
                const int oi = orig[ia[k]];
                R.ex_lo[k] = excl_start[oi];
                R.ex_hi[k] = excl_start[oi + 1];
            }
        }
    }
    uint32_t n_in[TILE_A], n_out[TILE_A];
#pragma unroll
    for (int k = 0; k < TILE_A; ++k) n_in[k] = n_out[k] = 0u;
    for (uint32_t t0 = 0; t0 < M.m;) {
        const uint32_t B = chunk_B(M.m - t0);
        sweep_chunk<NA, WRAP, PART>(tile, t0, B, g, rl2, rc2_inner, R.pi, R.t_self, lane, R.hit, R.inn);
        apply_exclusions<NA>(tile_slot, t0 + (uint32_t)lane * B, R.ex_lo, R.ex_hi, orig, excl_idx, R.hit);
#pragma unroll
        for (int k = 0; k < NA; ++k) {
            if (PART) {
                n_in[k] += __popc(R.hit[k] & R.inn[k]);
                n_out[k] += __popc(R.hit[k] & ~R.inn[k]);
            } else {
                n_in[k] += __popc(R.hit[k]);
            }
        }
        t0 += 32u * B;
    }
#pragma unroll
    for (int k = 0; k < NA; ++k) {
        uint32_t tot_in, tot_out = 0u;
        R.lane_in[k] = warp_excl_scan(n_in[k], lane, &tot_in);
        R.lane_out[k] = 0u;
        if (PART) R.lane_out[k] = warp_excl_scan(n_out[k], lane, &tot_out);
        R.len[k] = tot_in + tot_out;
        R.n_inner[k] = tot_in;
    }
}

// phase 2: write the rows in tile order; inner entries from the front, skin-shell entries from the back
// IDX = uint32_t: entries are global slots (tile_slot[t]); IDX = uint16_t: entries are the TILE-LOCAL indices t themselves
// (the compact list pair_tile.cu gathers from its own copy of the tile, laid out by the same tile_plan)
template <int NA, int WRAP, bool PART, typename IDX>
__device__ __forceinline__ void rows_phase2(const float4 *tile, const uint32_t *tile_slot, const StageMeta &M, const GridParams &g,
                                            float rl2, float rc2_inner, const uint32_t (&row)[TILE_A],
                                            const int *__restrict__ orig, const int32_t *__restrict__ excl_idx,
                                            IDX *__restrict__ nbr_list, int lane, RowState &R) {
    const bool one_chunk = M.m <= 992u;
    uint32_t run_in[TILE_A], run_out[TILE_A];
#pragma unroll
    for (int k = 0; k < TILE_A; ++k) run_in[k] = run_out[k] = 0u;
    for (uint32_t t0 = 0; t0 < M.m;) {
        const uint32_t B = chunk_B(M.m - t0);
        const uint32_t tl = t0 + (uint32_t)lane * B;
        if (!one_chunk) {
            sweep_chunk<NA, WRAP, PART>(tile, t0, B, g, rl2, rc2_inner, R.pi, R.t_self, lane, R.hit, R.inn);
            apply_exclusions<NA>(tile_slot, tl, R.ex_lo, R.ex_hi, orig, excl_idx, R.hit);
        }
#pragma unroll
        for (int k = 0; k < NA; ++k) {
            uint32_t li = R.lane_in[k], lo = R.lane_out[k];
            if (!one_chunk) {  // per-chunk lane offsets
                uint32_t ti, to = 0u;
                li = warp_excl_scan(PART ? __popc(R.hit[k] & R.inn[k]) : __popc(R.hit[k]), lane, &ti);
                lo = 0u;
                if (PART) lo = warp_excl_scan(__popc(R.hit[k] & ~R.inn[k]), lane, &to);
                li += run_in[k]; lo += run_out[k];
                run_in[k] += ti; run_out[k] += to;
            }
            uint32_t p_in = row[k] + li;
            uint32_t p_out = row[k] + R.len[k] - 1u - lo;
            uint32_t mleft = R.hit[k];
            while (mleft) {
                const int it = __ffs(mleft) - 1;
                mleft &= mleft - 1u;
                const IDX j = sizeof(IDX) == 2 ? (IDX)(tl + (uint32_t)it) : (IDX)tile_slot[tl + (uint32_t)it];
                if (!PART || ((R.inn[k] >> it) & 1u)) nbr_list[p_in++] = j;
                else nbr_list[p_out--] = j;
            }
        }
        t0 += 32u * B;
    }
}

// Minimum resident CTAs per SM the register allocation is held to: 3 -> 72 registers and 296 B of spill stores, 2 -> 96
// registers and 40 B (ptxas -v, sm_100a).  Which one is faster is a measurement: -DMC_TILE_MIN_BLOCKS=2 builds the variant.
#ifndef MC_TILE_MIN_BLOCKS
#define MC_TILE_MIN_BLOCKS 3
#endif
template <typename IDX, bool PART>
__global__ void __launch_bounds__((TILE_WARPS + 1) * 32, MC_TILE_MIN_BLOCKS) tile_build_kernel(
    int n_rows, const float4 *__restrict__ xyzq, const uint32_t *__restrict__ cell_start,
    const GridParams *__restrict__ gp, float rl2, float rc2_inner, const int *__restrict__ orig,
    const int32_t *__restrict__ excl_start, const int32_t *__restrict__ excl_idx, uint32_t *__restrict__ nbr_count,
    uint32_t *__restrict__ nbr_start, IDX *__restrict__ nbr_list, uint32_t list_cap, uint32_t tile_cap, int split,
    int n_stages /* 1 or 2 tiles in flight */, uint32_t *__restrict__ ctl /* [0] work counter, [1] list cursor, [2] max tile atoms seen, [3] tile overflow, [4] largest row block (entries) of one pass, [5] most atoms in a row cell */) {
    MC_DYN_SHARED_ALIGNED(unsigned char, smem_raw, 128);
    // per stage: tile_cap float4 positions, then tile_cap slot ids
    const size_t stage_bytes = (size_t)tile_cap * (sizeof(float4) + sizeof(uint32_t));
    __shared__ __align__(8) uint64_t full_bar[TILE_STAGES], empty_bar[TILE_STAGES];
    __shared__ StageMeta meta[TILE_STAGES];
    __shared__ uint32_t s_len[2][TILE_WARPS * TILE_A], s_off[2][TILE_WARPS * TILE_A];
    __shared__ int s_fits[2];

    const GridParams g = *gp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TILE_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], TILE_WARPS);
        }
    }
    __syncthreads();
    const long long n_items = (long long)g.ncell * split;

    if (warp == 0) {
        // ===== producer =====
        uint32_t it = 0;
        for (;;) {
            long long w = 0;
            if (lane == 0) w = (long long)atomicAdd(ctl, 1u);  // dynamic distribution: boundary cells cost more
            w = __shfl_sync(MC_FULL_MASK, w, 0);
            const bool done = w >= n_items;
            const int c = done ? 0 : (int)(w / split);
            uint32_t a0 = 0xffffffffu, a1 = 0xffffffffu, m = 0, self_off = 0;
            TileRange r0 = {0u, 0u, 0u}, r1 = {0u, 0u, 0u};
            int wrap = 0;
            if (!done) {
                a0 = cell_start[c];
                a1 = cell_start[c + 1];
                if (a0 == a1) continue;  // empty cell: nothing to stage (warp-uniform)
                const int c2 = c / (g.nc[0] * g.nc[1]);
                if (c2 < g.row_l0 || c2 >= g.row_l1) continue;  // ghost layer: its atoms carry no rows
                TilePlan P;
                tile_plan(g, cell_start, c, a0, lane, P);  // tile layout shared with the force kernel (tile_ring.cuh)
                r0 = P.r0; r1 = P.r1; m = P.m; self_off = P.self_off; wrap = P.wrap;
                if (lane == 0) { atomicMax(ctl + 2, m); if (a1 - a0 > *reinterpret_cast<volatile uint32_t *>(ctl + 5)) atomicMax(ctl + 5, a1 - a0); }
                if (padded_size(m) > tile_cap) {  // does not fit: the host enlarges the tile (or falls back)
                    if (lane == 0) ctl[3] = 1u;
                    continue;
                }
            }
            const int s = it % n_stages;
            mbar_wait(&empty_bar[s], ((it / n_stages) & 1) ^ 1);
            float4 *tile = reinterpret_cast<float4 *>(smem_raw + (size_t)s * stage_bytes);
            uint32_t *tile_slot = reinterpret_cast<uint32_t *>(tile + tile_cap);
            if (lane == 0) {
                meta[s].m = m; meta[s].a0 = a0; meta[s].a1 = a1; meta[s].wrap = wrap; meta[s].self_off = self_off;
                meta[s].slice = done ? 0u : (uint32_t)(w % split);
            }
            // slot ids of the staged atoms (lanes 0..8 own the ranges; every lane helps to write them)
            for (int src_lane = 0; src_lane < 9; ++src_lane) {
                const uint32_t s0 = __shfl_sync(MC_FULL_MASK, r0.src, src_lane), n0 = __shfl_sync(MC_FULL_MASK, r0.cnt, src_lane),
                               o0 = __shfl_sync(MC_FULL_MASK, r0.off, src_lane), s1 = __shfl_sync(MC_FULL_MASK, r1.src, src_lane),
                               n1 = __shfl_sync(MC_FULL_MASK, r1.cnt, src_lane), o1 = __shfl_sync(MC_FULL_MASK, r1.off, src_lane);
                for (uint32_t t = lane; t < n0; t += 32) tile_slot[o0 + t] = s0 + t;
                for (uint32_t t = lane; t < n1; t += 32) tile_slot[o1 + t] = s1 + t;
            }
            // pad the tile to whole chunks with NaN positions: the sweep needs no bounds test
            {
                const float qnan = __int_as_float(0x7fffffff);
                const uint32_t pend = padded_size(m);
                for (uint32_t t = m + lane; t < pend; t += 32) tile[t] = make_float4(qnan, qnan, qnan, 0.f);
            }
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&full_bar[s], m * (uint32_t)sizeof(float4));  // release: meta + slots visible
            __syncwarp();
            if (lane < 9) {
                if (r0.cnt) tma_bulk_g2s(tile + r0.off, xyzq + r0.src, r0.cnt * (uint32_t)sizeof(float4), &full_bar[s]);
                if (r1.cnt) tma_bulk_g2s(tile + r1.off, xyzq + r1.src, r1.cnt * (uint32_t)sizeof(float4), &full_bar[s]);
            }
            ++it;
            if (done) break;
        }
    } else {
        // ===== consumers =====
        const int cw = warp - 1;
        int pp = 0;
        for (uint32_t it = 0;; ++it) {
            const int s = it % n_stages;
            mbar_wait(&full_bar[s], (it / n_stages) & 1);
            const StageMeta M = meta[s];
            if (M.a0 == 0xffffffffu) break;
            const float4 *tile = reinterpret_cast<const float4 *>(smem_raw + (size_t)s * stage_bytes);
            const uint32_t *tile_slot = reinterpret_cast<const uint32_t *>(tile + tile_cap);
            for (uint32_t base = M.a0 + M.slice * (TILE_WARPS * TILE_A); base < M.a1; base += TILE_WARPS * TILE_A * split) {
                uint32_t ia[TILE_A];
                int na = 0;
#pragma unroll
                for (int k = 0; k < TILE_A; ++k) {
                    const uint32_t i = base + cw + k * TILE_WARPS;  // round-robin: every warp gets 2-3 atoms of a ~20-atom cell
                    ia[k] = i;
                    if (i < M.a1 && (int)i < n_rows) na = k + 1;
                }
                // every warp of the CTA runs the SAME instantiation (atoms in this pass / 8, rounded up): one hot
                // loop in the instruction cache; surplus slots are parked on NaN positions
                const uint32_t n_pass = min((uint32_t)(TILE_WARPS * TILE_A), M.a1 - base);
                const int na_u = (int)((n_pass + TILE_WARPS - 1) / TILE_WARPS);
                RowState R;
#define MC_P1(NA_, W_) rows_phase1<NA_, W_, PART>(tile, tile_slot, M, g, rl2, rc2_inner, ia, xyzq, orig, excl_start, excl_idx, lane, na, R)
#define MC_P1W(NA_) \
    if (M.wrap == 0) MC_P1(NA_, 0); else if (M.wrap == 1) MC_P1(NA_, 1); else MC_P1(NA_, 2)
                // whole pairs of atoms (packed arithmetic): two instantiations per wrap class
                if (na_u <= 2) { MC_P1W(2); } else { MC_P1W(4); }
#undef MC_P1W
#undef MC_P1
                // ---- row allocation for the whole pass (up to 32 atoms of the cell, in atom order): the padded
                // lengths go through shared memory, consumer warp 0 scans them and claims the space with ONE
                // atomicAdd, so the rows of a cell are contiguous and ordered like its atoms
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < TILE_A; ++k) s_len[pp][cw + k * TILE_WARPS] = k < na ? ((R.len[k] + 7u) & ~7u) : 0u;
                }
                consumer_sync();
                if (cw == 0) {
                    uint32_t tot;
                    const uint32_t off = warp_excl_scan(s_len[pp][lane], lane, &tot);
                    uint32_t b0 = 0;
                    if (lane == 0) { b0 = atomicAdd(ctl + 1, tot); if (tot > *reinterpret_cast<volatile uint32_t *>(ctl + 4)) atomicMax(ctl + 4, tot); }
                    b0 = __shfl_sync(MC_FULL_MASK, b0, 0);
                    s_off[pp][lane] = b0 + off;
                    if (lane == 0) s_fits[pp] = ((uint64_t)b0 + tot <= (uint64_t)list_cap) ? 1 : 0;
                }
                consumer_sync();
                uint32_t row[TILE_A];
#pragma unroll
                for (int k = 0; k < TILE_A; ++k) {
                    row[k] = s_off[pp][cw + k * TILE_WARPS];
                    if (k < na && lane == 0) {
                        nbr_start[ia[k]] = row[k];
                        nbr_count[ia[k]] = R.len[k];
                    }
                }
                const bool fits = s_fits[pp] != 0;  // otherwise the host grows the list and rebuilds
                pp ^= 1;  // the other buffer serves the next pass: no third barrier needed
                if (fits && na > 0) {
#define MC_P2(NA_, W_) rows_phase2<NA_, W_, PART, IDX>(tile, tile_slot, M, g, rl2, rc2_inner, row, orig, excl_idx, nbr_list, lane, R)
#define MC_P2W(NA_) \
    if (M.wrap == 0) MC_P2(NA_, 0); else if (M.wrap == 1) MC_P2(NA_, 1); else MC_P2(NA_, 2)
                    if (na_u <= 2) { MC_P2W(2); } else { MC_P2W(4); }
#undef MC_P2W
#undef MC_P2
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// rows_build_kernel -- the default build (global-slot rows, no inner / skin partition).  Same producer, same tile, same
// accept expression; the consumer side is organised around what the first kernel's profile showed (round 2, ncu: only a
// third of its 725 M warp instructions were distance tests -- 23 % were per-lane find-first-set write loops running at
// 5 of 32 lanes, 40 % of its atom slots were parked because the eight warps of a CTA went through a cell in lock-step):
//   * lanes own CONSECUTIVE candidates (t = 32 b + lane): a block of 32 tile entries is one conflict-free LDS.128 per
//     lane, the warp-wide accept decision of an atom is one ballot, and tile order is ballot order -- a hit's place in its
//     row is count_so_far + popc(ballot & lanes_below), no per-lane loop, no prefix scan, no bit reversal;
//   * a warp sweeps the tile ONCE for a quad of four consecutive atoms of the cell, staging the rows as 16-bit tile
//     indices in shared memory; when the sweep ends the row lengths are known, the quad's rows are claimed with one
//     atomicAdd on the list cursor and copied out with coalesced 128-byte stores (slot ids looked up on the way).  A row
//     longer than the staging space is written by a second sweep straight to its place (dense systems: every row);
//   * the consumer warps are decoupled: quads are dealt round-robin ACROSS items (the deal continues where the previous
//     cell stopped), a warp that has no quad in this tile moves on to the next stage of the ring, so nobody is parked and
//     nobody waits at a CTA barrier; packing pairs the x / y components of ONE atom (FADD2 / FMUL2 on the register pair the
//     LDS.128 delivers), so any number of atoms per quad is as cheap per atom and nothing is duplicated into pairs.
// Row content and order are those of tile_build_kernel (ascending tile index); row placement follows completion order.
constexpr int RB_A = 4;           // atoms per quad
constexpr int RB_MAX_STAGES = 4;  // tiles in flight per CTA
constexpr int RB_WARPS_DENSE = 16;  // consumer warps of the one-CTA-per-SM configuration (dense systems)
constexpr size_t RB_DENSE_BUDGET = 226u * 1024u;  // dynamic shared memory of that configuration

__device__ __forceinline__ uint32_t lanemask_lt(int lane) { return (1u << lane) - 1u; }

// staging stores through a 32-bit shared-window address (one IADD3 forms base + 2 * rank; a generic pointer costs a 64-bit add)
#ifndef MC_HOST_SHIM
typedef uint32_t rb_sptr;
__device__ __forceinline__ rb_sptr rb_sptr_of(const uint16_t *p) { return smem_u32(p); }
__device__ __forceinline__ void rb_store16(rb_sptr a, uint16_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(v) : "memory"); }
#else
typedef uintptr_t rb_sptr;
inline rb_sptr rb_sptr_of(const uint16_t *p) { return reinterpret_cast<uintptr_t>(p); }
inline void rb_store16(rb_sptr a, uint16_t v) { *reinterpret_cast<uint16_t *>(a) = v; }
#endif

// accept decision of one atom (negated position npi = -x_i) against the candidate pj: the oracle's
// ((dx*dx)+(dy*dy))+(dz*dz) < rl2 on d = x_j - x_i = -(x_i - x_j) (exactly the negative in fp32, and every step below is
// odd or even in d: the decision is the oracle's bit for bit).  NaN (tile padding, parked atoms) is never accepted.
template <int WRAP>
__device__ __forceinline__ bool rb_accept(const float4 pj, const float2 npi_xy, const float npi_z, const GridParams &g, float rl2) {
    float2 dxy = mc_add2(make_float2(pj.x, pj.y), npi_xy);
    float dz = __fadd_rn(pj.z, npi_z);
    if (WRAP == 2) {
        dxy.x = min_image_exact(dxy.x, g.ext[0], g.inv_ext[0]);
        dxy.y = min_image_exact(dxy.y, g.ext[1], g.inv_ext[1]);
        dz = min_image_exact(dz, g.ext[2], g.inv_ext[2]);
    } else if (WRAP == 1) {  // see sweep_chunk: >= 3 cells per axis, the fast image is the exact one wherever it matters
        const float2 q = mc_mul2(dxy, make_float2(g.inv_ext[0], g.inv_ext[1]));
        dxy = mc_fma2(make_float2(-rintf(q.x), -rintf(q.y)), make_float2(g.ext[0], g.ext[1]), dxy);
        dz = __fmaf_rn(-rintf(__fmul_rn(dz, g.inv_ext[2])), g.ext[2], dz);
    }
    const float2 s = mc_mul2(dxy, dxy);
    const float r2 = __fadd_rn(__fadd_rn(s.x, s.y), __fmul_rn(dz, dz));
    return r2 < rl2;
}

struct RbQuad {
    float2 npxy[RB_A];
    float npz[RB_A];
    uint32_t t_self[RB_A];
    int ex_lo[RB_A], ex_hi[RB_A];
};

// One pass of a quad over the tile.  MODE 0: stage the rows (16-bit tile indices, the first `stage_cap` entries of each)
// and count; MODE 1: write the rows flagged in `direct` straight to nbr_list at row[k] (second pass of rows that did not
// fit the staging space); MODE 2: ONE pass that writes every row straight to the space claimed for it beforehand (at most
// `stage_cap` entries each -- the argument carries the row capacity here -- and counts on, so that an overflow is seen).  cnt[k] = row lengths (warp-uniform).  EXCL: some atom of the quad carries exclusions (rare; the
// common instantiation keeps every decision in a predicate register).
template <int WRAP, int MODE, bool EXCL>
__device__ __forceinline__ void rb_sweep(const float4 *tile, const uint32_t *tile_slot, uint32_t m_pad, const GridParams &g, float rl2,
                                         const RbQuad &Q, const int *__restrict__ orig, const int32_t *__restrict__ excl_idx,
                                         uint16_t *stage, uint32_t stage_cap, const uint32_t (&row)[RB_A], uint32_t direct,
                                         uint32_t *__restrict__ nbr_list, int lane, uint32_t (&cnt)[RB_A]) {
    // MODE 0: byte address (shared window) of the next free staging element of row k and the end of its space;
    // MODE 1: index of the next list entry of row k
    // (the staged rows are interleaved: element e of row k sits at stage[e * RB_A + k], so one end-of-space bound serves
    // all rows and the row's offset is an immediate of the store)
    rb_sptr sp[RB_A];
    const rb_sptr se = rb_sptr_of(stage) + 2u * RB_A * stage_cap;
    uint32_t lo[RB_A];
#pragma unroll
    for (int k = 0; k < RB_A; ++k) {
        sp[k] = rb_sptr_of(stage) + 2u * (uint32_t)k;
        lo[k] = row[k];
    }
    const uint32_t lt = lanemask_lt(lane);
    for (uint32_t t = (uint32_t)lane; t < m_pad; t += 32u) {
        const float4 pj = tile[t];
#pragma unroll
        for (int k = 0; k < RB_A; ++k) {
            bool hit = rb_accept<WRAP>(pj, Q.npxy[k], Q.npz[k], g, rl2) && t != Q.t_self[k];
            if (EXCL) {  // 1-2 / 1-3 / 1-4 partners (original ids) never enter the list: only hits of atoms that carry exclusions pay
                if (hit && Q.ex_hi[k] > Q.ex_lo[k]) {
                    const int oj = orig[tile_slot[t]];
                    for (int e = Q.ex_lo[k]; e < Q.ex_hi[k]; ++e)
                        if (excl_idx[e] == oj) { hit = false; break; }
                }
            }
            const uint32_t mask = __ballot_sync(MC_FULL_MASK, hit);
            const uint32_t below = (uint32_t)__popc(mask & lt), all = (uint32_t)__popc(mask);
            if (MODE == 0) {
                const rb_sptr q = sp[k] + 2u * RB_A * below;
                if (hit && q < se) rb_store16(q, (uint16_t)t);
                sp[k] += 2u * RB_A * all;
            } else if (MODE == 1) {
                if (hit && ((direct >> k) & 1u)) nbr_list[lo[k] + below] = tile_slot[t];
                lo[k] += all;
            } else {
                if (hit && lo[k] + below - row[k] < stage_cap) nbr_list[lo[k] + below] = tile_slot[t];
                lo[k] += all;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < RB_A; ++k)
        cnt[k] = MODE == 0 ? (uint32_t)((sp[k] - rb_sptr_of(stage)) / (2u * RB_A)) : lo[k] - row[k];
}

// Exclusions of rows_build_kernel, after the sweep instead of inside it: the sweep lists every atom inside the radius (no
// per-hit gather of the candidate's original id -- a dependent global round trip that bounded the build of dense solvated
// systems), then the quad's rows are filtered in place: lane x < n_ex holds the SLOT of the atom's x-th excluded partner
// (slot_of_orig of its original id; -1 = not held by this rank), a row entry equal to one of them is dropped and the entries
// behind it move up.  Chunks of 32 entries: every lane has read its entry before the ballot, the kept ones are written at or
// below the chunk's own first position.  Returns the new row length.
__device__ __forceinline__ uint32_t rb_filter_row(uint32_t *__restrict__ nbr_list, uint32_t row, uint32_t n, uint32_t ps, int n_ex, int lane) {
    const uint32_t lt = lanemask_lt(lane);
    uint32_t w = 0u;
    for (uint32_t b = 0u; b < n; b += 32u) {
        const uint32_t e = b + (uint32_t)lane;
        bool keep = e < n;
        uint32_t v = 0xffffffffu;
        if (keep) v = nbr_list[row + e];
        for (int x = 0; x < n_ex; ++x) {
            const uint32_t partner = __shfl_sync(MC_FULL_MASK, ps, x);  // every lane takes part (no short-circuit around it)
            keep = keep && v != partner;
        }
        const uint32_t mask = __ballot_sync(MC_FULL_MASK, keep);
        if (keep) nbr_list[row + w + (uint32_t)__popc(mask & lt)] = v;
        w += (uint32_t)__popc(mask);
    }
    return w;
}

template <int WRAP, bool EXK /* the system carries exclusions (its own instantiation: the plain one keeps its 72 registers) */>
__device__ __forceinline__ void rb_quad(const float4 *tile, const uint32_t *tile_slot, const StageMeta &M, uint32_t m_pad, uint32_t i0,
                                        int n_rows, const GridParams &g, float rl2, const int *__restrict__ orig,
                                        const int32_t *__restrict__ excl_start, const int32_t *__restrict__ excl_idx, uint16_t *stage,
                                        uint32_t stage_cap, uint32_t *__restrict__ nbr_count, uint32_t *__restrict__ nbr_start,
                                        uint32_t *__restrict__ nbr_list, uint32_t list_cap, uint32_t *__restrict__ ctl, int lane,
                                        uint32_t row_cap, const int *__restrict__ slot_of_orig) {
    const float qnan = __int_as_float(0x7fffffff);
    RbQuad Q;
    uint32_t valid = 0u;
    bool any_excl = false;
#pragma unroll
    for (int k = 0; k < RB_A; ++k) {
        const uint32_t i = i0 + (uint32_t)k;
        Q.npxy[k] = make_float2(qnan, qnan);
        Q.npz[k] = qnan;
        Q.t_self[k] = 0xffffffffu;
        Q.ex_lo[k] = Q.ex_hi[k] = 0;
        if (i < M.a1 && (int)i < n_rows) {
            valid |= 1u << k;
            Q.t_self[k] = M.self_off + (i - M.a0);
            const float4 p = tile[Q.t_self[k]];  // the own cell is part of the staged tile
            Q.npxy[k] = make_float2(-p.x, -p.y);
            Q.npz[k] = -p.z;
            if (EXK && excl_start) {
                const int oi = orig[i];
                Q.ex_lo[k] = excl_start[oi];
                Q.ex_hi[k] = excl_start[oi + 1];
                any_excl = any_excl || Q.ex_hi[k] > Q.ex_lo[k];
            }
        }
    }
    if (!valid) return;
    uint32_t cnt[RB_A], row[RB_A] = {0u, 0u, 0u, 0u};
    if (!EXK) any_excl = false;
    // exclusions: filtered after the sweep (rb_filter_row) when every atom of the quad has at most 32 partners and the slot
    // table is there; else inside the sweep (EXCL instantiation)
    bool post = EXK && any_excl && slot_of_orig != nullptr;
    uint32_t ps[RB_A];
#pragma unroll
    for (int k = 0; k < RB_A; ++k) {
        ps[k] = 0xffffffffu;
        if (Q.ex_hi[k] - Q.ex_lo[k] > 32) post = false;
    }
    if (post) {
#pragma unroll
        for (int k = 0; k < RB_A; ++k)
            if (lane < Q.ex_hi[k] - Q.ex_lo[k]) ps[k] = (uint32_t)slot_of_orig[excl_idx[Q.ex_lo[k] + lane]];
    }
    const bool sweep_excl = EXK && any_excl && !post;
    if (row_cap) {
        // Dense systems (rows of a thousand entries: no room to stage them next to a 100+ KB tile): the quad claims row_cap
        // entries per row up front -- the longest row of the previous build + 25 % -- and writes them in ONE sweep instead of
        // counting first and sweeping again.  A row that outgrows its space raises bit 1 of ctl[3]; the host then builds again
        // with the longer hint.  Rows keep their own counts, the padding is never read.
        uint32_t base = 0u;
        if (lane == 0) base = atomicAdd(ctl + 1, RB_A * row_cap);
        base = __shfl_sync(MC_FULL_MASK, base, 0);
        if ((uint64_t)base + RB_A * row_cap > (uint64_t)list_cap) return;  // the host grows the list and builds again
#pragma unroll
        for (int k = 0; k < RB_A; ++k) row[k] = base + (uint32_t)k * row_cap;
        if (EXK && sweep_excl) rb_sweep<WRAP, 2, EXK>(tile, tile_slot, m_pad, g, rl2, Q, orig, excl_idx, stage, row_cap, row, 0u, nbr_list, lane, cnt);
        else rb_sweep<WRAP, 2, false>(tile, tile_slot, m_pad, g, rl2, Q, orig, excl_idx, stage, row_cap, row, 0u, nbr_list, lane, cnt);
        uint32_t mx = 0u;
        if (post) __syncwarp();  // the rows written by all lanes are visible to the warp
#pragma unroll
        for (int k = 0; k < RB_A; ++k) {
            mx = max(mx, cnt[k]);  // (before the filter: the space is claimed for the unfiltered row)
            uint32_t len = min(cnt[k], row_cap);
            if (post && Q.ex_hi[k] > Q.ex_lo[k]) len = rb_filter_row(nbr_list, row[k], len, ps[k], Q.ex_hi[k] - Q.ex_lo[k], lane);
            if (lane == k && ((valid >> k) & 1u)) {
                nbr_start[i0 + (uint32_t)k] = row[k];
                nbr_count[i0 + (uint32_t)k] = len;
            }
        }
        if (lane == 0) {
            if (mx > *reinterpret_cast<volatile uint32_t *>(ctl + 6)) atomicMax(ctl + 6, mx);
            if (mx > row_cap) atomicOr(ctl + 3, 2u);
        }
        return;
    }
    if (EXK && sweep_excl) rb_sweep<WRAP, 0, EXK>(tile, tile_slot, m_pad, g, rl2, Q, orig, excl_idx, stage, stage_cap, row, 0u, nbr_list, lane, cnt);
    else rb_sweep<WRAP, 0, false>(tile, tile_slot, m_pad, g, rl2, Q, orig, excl_idx, stage, stage_cap, row, 0u, nbr_list, lane, cnt);
    // claim the quad's rows (each padded to 8 entries = whole 32-byte sectors) with one atomicAdd
    uint32_t tot = 0u, mx = 0u, direct = 0u;
#pragma unroll
    for (int k = 0; k < RB_A; ++k) {
        row[k] = tot;
        tot += (cnt[k] + 7u) & ~7u;
        mx = max(mx, cnt[k]);
        if (cnt[k] > stage_cap) direct |= 1u << k;
    }
    uint32_t base = 0u;
    if (lane == 0) {
        base = atomicAdd(ctl + 1, tot);
        if (mx > *reinterpret_cast<volatile uint32_t *>(ctl + 6)) atomicMax(ctl + 6, mx);
    }
    base = __shfl_sync(MC_FULL_MASK, base, 0);
#pragma unroll
    for (int k = 0; k < RB_A; ++k) {
        row[k] += base;
        if (lane == k && ((valid >> k) & 1u)) {
            nbr_start[i0 + (uint32_t)k] = row[k];
            nbr_count[i0 + (uint32_t)k] = cnt[k];
        }
    }
    if ((uint64_t)base + tot > (uint64_t)list_cap) return;  // the host grows the list and builds again
    __syncwarp();  // the staged entries of every lane are visible to the warp
#pragma unroll
    for (int k = 0; k < RB_A; ++k) {
        if ((direct >> k) & 1u) continue;
        for (uint32_t e = (uint32_t)lane; e < cnt[k]; e += 32u) nbr_list[row[k] + e] = tile_slot[stage[e * RB_A + (uint32_t)k]];
    }
    if (direct) {
        uint32_t cnt2[RB_A];
        if (EXK && sweep_excl) rb_sweep<WRAP, 1, EXK>(tile, tile_slot, m_pad, g, rl2, Q, orig, excl_idx, stage, stage_cap, row, direct, nbr_list, lane, cnt2);
        else rb_sweep<WRAP, 1, false>(tile, tile_slot, m_pad, g, rl2, Q, orig, excl_idx, stage, stage_cap, row, direct, nbr_list, lane, cnt2);
    }
    __syncwarp();  // copy-out reads done before the next quad's staging overwrites the space (and the rows are visible to the warp)
    if (post) {
#pragma unroll
        for (int k = 0; k < RB_A; ++k) {
            if (Q.ex_hi[k] <= Q.ex_lo[k]) continue;  // warp-uniform
            const uint32_t len = rb_filter_row(nbr_list, row[k], cnt[k], ps[k], Q.ex_hi[k] - Q.ex_lo[k], lane);
            if (lane == k && ((valid >> k) & 1u)) nbr_count[i0 + (uint32_t)k] = len;
        }
    }
}

// Per-cell staging record of rows_build_kernel's producer (RB_PLAN_WORDS words, written once per build by
// rows_plan_kernel): [0] a0 [1] a1 [2] m [3] self_off [4] wrap | skip << 8; then src[18], cnt[18], off[18] of the tile's
// ranges.  With the record the producer needs ONE round trip to global memory per item instead of three dependent ones
// (work counter -> cell_start -> stencil rows), and that one is issued an item ahead.
constexpr int RB_PLAN_WORDS = 64;
constexpr int RB_PLAN_SRC = 8, RB_PLAN_CNT = 26, RB_PLAN_OFF = 44;

__global__ void __launch_bounds__(256) rows_plan_kernel(const uint32_t *__restrict__ cell_start, const GridParams *__restrict__ gp,
                                                        uint32_t *__restrict__ plan, uint32_t *__restrict__ ctl) {
    const GridParams g = *gp;
    const int lane = threadIdx.x & 31;
    const int c = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (c >= g.ncell) return;  // warp-uniform
    const uint32_t a0 = cell_start[c], a1 = cell_start[c + 1];
    const int c2 = c / (g.nc[0] * g.nc[1]);
    const bool skip = a0 == a1 || c2 < g.row_l0 || c2 >= g.row_l1;  // empty cell / ghost layer: no rows
    uint32_t *rec = plan + (size_t)c * RB_PLAN_WORDS;
    TilePlan P;
    P.m = 0; P.self_off = 0; P.wrap = 0;
    P.r0 = TileRange{0u, 0u, 0u}; P.r1 = TileRange{0u, 0u, 0u};
    if (!skip) tile_plan(g, cell_start, c, a0, lane, P);
    if (lane < 9) {
        rec[RB_PLAN_SRC + 2 * lane] = P.r0.src; rec[RB_PLAN_CNT + 2 * lane] = P.r0.cnt; rec[RB_PLAN_OFF + 2 * lane] = P.r0.off;
        rec[RB_PLAN_SRC + 2 * lane + 1] = P.r1.src; rec[RB_PLAN_CNT + 2 * lane + 1] = P.r1.cnt; rec[RB_PLAN_OFF + 2 * lane + 1] = P.r1.off;
    }
    if (lane == 0) {
        rec[0] = a0; rec[1] = a1; rec[2] = P.m; rec[3] = P.self_off; rec[4] = (uint32_t)P.wrap | (skip ? 0x100u : 0u);
        if (!skip) {
            if (P.m > *reinterpret_cast<volatile uint32_t *>(ctl + 2)) atomicMax(ctl + 2, P.m);
            if (a1 - a0 > *reinterpret_cast<volatile uint32_t *>(ctl + 5)) atomicMax(ctl + 5, a1 - a0);
        }
    }
}

struct RbPlanRegs { uint32_t a0, a1, m, self_off, flags, src, cnt, off; };

__device__ __forceinline__ void rb_load_plan(const uint32_t *__restrict__ plan, int c, int lane, RbPlanRegs &R) {
    const uint32_t *rec = plan + (size_t)c * RB_PLAN_WORDS;
    R.a0 = rec[0]; R.a1 = rec[1]; R.m = rec[2]; R.self_off = rec[3]; R.flags = rec[4];
    const int l = lane < 18 ? lane : 0;
    R.src = rec[RB_PLAN_SRC + l]; R.cnt = lane < 18 ? rec[RB_PLAN_CNT + l] : 0u; R.off = rec[RB_PLAN_OFF + l];
}

#ifndef MC_HOST_SHIM
// mbarrier wait that gives the issue slots away while it waits (a plain try_wait loop measured 15 % of the kernel's issued
// instructions: consumer warps run ahead of the tiles by design)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity) {
    uint32_t ok, ns = 250u;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(ns);
        if (ns < 2000u) ns += ns;
    }
}
#else
inline void mbar_wait_sleep(uint64_t *bar, uint32_t parity) { shim_mbar_wait(bar, parity); }
#endif

// NW consumer warps: 8 where several CTAs share an SM; 16 where the tile leaves room for one CTA only (dense systems), which
// would otherwise run 9 warps per SM and wait on every shared-memory round trip.
template <int MINB, int NW, bool EXK>
__global__ void __launch_bounds__((NW + 1) * 32, MINB) rows_build_kernel(
    int n_rows, const float4 *__restrict__ xyzq, const uint32_t *__restrict__ plan, const GridParams *__restrict__ gp, float rl2,
    const int *__restrict__ orig, const int32_t *__restrict__ excl_start, const int32_t *__restrict__ excl_idx,
    uint32_t *__restrict__ nbr_count, uint32_t *__restrict__ nbr_start, uint32_t *__restrict__ nbr_list, uint32_t list_cap,
    uint32_t tile_cap, uint32_t stage_cap /* staged entries per row; 0: every row takes two sweeps */, int split, int n_stages,
    uint32_t *__restrict__ ctl /* as tile_build_kernel; [3] bit 1: a row outgrew row_cap; [6] longest row */,
    uint32_t row_cap /* > 0 (with stage_cap == 0): single direct sweep into row_cap entries per row */,
    const int *__restrict__ slot_of_orig /* exclusions are filtered after the sweep when given (rb_filter_row) */) {
    MC_DYN_SHARED_ALIGNED(unsigned char, smem_raw, 128);
    const size_t stage_bytes = (size_t)tile_cap * (sizeof(float4) + sizeof(uint32_t));
    __shared__ __align__(8) uint64_t full_bar[RB_MAX_STAGES], empty_bar[RB_MAX_STAGES];
    __shared__ StageMeta meta[RB_MAX_STAGES];

    const GridParams g = *gp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < RB_MAX_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NW);
        }
    }
    __syncthreads();
    const long long n_items = (long long)g.ncell * split;

    if (warp == 0) {
        // ===== producer =====
        // Items come from the global work counter (cells differ in cost -- the minimum image of boundary cells -- and a
        // static deal aliases with the grid: 444 CTAs over rows of 37 cells gave every CTA one x coordinate).  Two things
        // are in flight while an item is staged: the plan record of the next item and the claim of the one after it.
        int s = 0;
        uint32_t ph = 1;  // parity to wait for on empty_bar[s] (flips every time the ring wraps)
        uint32_t pend = 0;
        if (lane == 0) pend = atomicAdd(ctl, 1u);
        long long w = (long long)__shfl_sync(MC_FULL_MASK, pend, 0);
        bool have = w < n_items;
        RbPlanRegs nx = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        uint32_t slice_nx = 0;
        if (have) { rb_load_plan(plan, (int)(w / split), lane, nx); slice_nx = (uint32_t)(w % split); }
        if (lane == 0 && have) pend = atomicAdd(ctl, 1u);
        for (;;) {
            const bool done = !have;
            const RbPlanRegs cu = nx;
            const uint32_t slice = slice_nx;
            if (!done) {
                w = (long long)__shfl_sync(MC_FULL_MASK, pend, 0);  // the claim issued an item ago
                have = w < n_items;
                if (have) { rb_load_plan(plan, (int)(w / split), lane, nx); slice_nx = (uint32_t)(w % split); }
                if (lane == 0 && have) pend = atomicAdd(ctl, 1u);
            }
            uint32_t a0 = 0xffffffffu, a1 = 0xffffffffu, m = 0, self_off = 0, rsrc = 0, rcnt = 0, roff = 0;
            int wrap = 0;
            if (!done) {
                if (cu.flags & 0x100u) continue;  // empty cell / ghost layer
                if (((cu.m + 31u) & ~31u) > tile_cap) {  // does not fit: the host enlarges the tile (or falls back)
                    if (lane == 0) atomicOr(ctl + 3, 1u);
                    continue;
                }
                a0 = cu.a0; a1 = cu.a1; m = cu.m; self_off = cu.self_off; wrap = (int)(cu.flags & 0xffu);
                rsrc = cu.src; rcnt = cu.cnt; roff = cu.off;
            }
            mbar_wait_sleep(&empty_bar[s], ph);
            float4 *tile = reinterpret_cast<float4 *>(smem_raw + (size_t)s * stage_bytes);
            uint32_t *tile_slot = reinterpret_cast<uint32_t *>(tile + tile_cap);
            if (lane == 0) {
                meta[s].m = m; meta[s].a0 = a0; meta[s].a1 = a1; meta[s].wrap = wrap; meta[s].self_off = self_off;
                meta[s].slice = slice;
            }
            // slot ids of the staged atoms (lanes 0..17 own a range each; every lane helps to write them)
            for (int src_lane = 0; src_lane < 18; ++src_lane) {
                const uint32_t n0 = __shfl_sync(MC_FULL_MASK, rcnt, src_lane);
                if (n0 == 0u) continue;  // warp-uniform
                const uint32_t s0 = __shfl_sync(MC_FULL_MASK, rsrc, src_lane), o0 = __shfl_sync(MC_FULL_MASK, roff, src_lane);
                for (uint32_t t = lane; t < n0; t += 32) tile_slot[o0 + t] = s0 + t;
            }
            {
                const float qnan = __int_as_float(0x7fffffff);
                const uint32_t pend = (m + 31u) & ~31u;
                for (uint32_t t = m + lane; t < pend; t += 32) tile[t] = make_float4(qnan, qnan, qnan, 0.f);
            }
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&full_bar[s], m * (uint32_t)sizeof(float4));  // release: meta + slots visible
            __syncwarp();
            if (rcnt) tma_bulk_g2s(tile + roff, xyzq + rsrc, rcnt * (uint32_t)sizeof(float4), &full_bar[s]);
            if (++s == n_stages) { s = 0; ph ^= 1u; }
            if (done) break;
        }
    } else {
        // ===== consumers: decoupled warps, quads dealt round-robin across items =====
        const int cw = warp - 1;
        uint16_t *stage = reinterpret_cast<uint16_t *>(smem_raw + (size_t)n_stages * stage_bytes) + (size_t)cw * RB_A * stage_cap;
        uint32_t rot = 0;  // quads dealt so far (mod NW): the same number on every warp
        int s = 0;
        uint32_t ph = 0;
        for (;;) {
            mbar_wait_sleep(&full_bar[s], ph);
            const StageMeta M = meta[s];
            if (M.a0 == 0xffffffffu) break;
            const float4 *tile = reinterpret_cast<const float4 *>(smem_raw + (size_t)s * stage_bytes);
            const uint32_t *tile_slot = reinterpret_cast<const uint32_t *>(tile + tile_cap);
            const uint32_t m_pad = (M.m + 31u) & ~31u;
            const uint32_t lim = min(M.a1, (uint32_t)max(n_rows, 0));
            const uint32_t n_at = lim > M.a0 ? lim - M.a0 : 0u;
            const uint32_t n_quads_cell = (n_at + RB_A - 1) / RB_A;
            // this item's quads: q = slice, slice + split, ...
            uint32_t nq = n_quads_cell;
            if (split > 1) nq = n_quads_cell > M.slice ? (n_quads_cell - M.slice + (uint32_t)split - 1u) / (uint32_t)split : 0u;
            for (uint32_t j = ((uint32_t)cw - rot) & (NW - 1); j < nq; j += NW) {
                const uint32_t i0 = M.a0 + (M.slice + j * (uint32_t)split) * RB_A;
#define MC_RBQ(W_) rb_quad<W_, EXK>(tile, tile_slot, M, m_pad, i0, n_rows, g, rl2, orig, excl_start, excl_idx, stage, stage_cap, nbr_count, \
                               nbr_start, nbr_list, list_cap, ctl, lane, row_cap, slot_of_orig)
                if (M.wrap == 0) MC_RBQ(0); else if (M.wrap == 1) MC_RBQ(1); else MC_RBQ(2);
#undef MC_RBQ
            }
            rot = (rot + nq) & (NW - 1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
    }
}
static_assert((TILE_WARPS & (TILE_WARPS - 1)) == 0, "the quad deal masks with TILE_WARPS - 1");

// Compact rows (tile-local 16-bit indices) -> rows of global slots, same nbr_start / nbr_count.  Off the hot path: only
// mc_get_neighbors, the virial and the between-molecules energy read global-slot rows.  One CTA per cell at a time.
__global__ void __launch_bounds__(128) expand_rows_kernel(const uint32_t *__restrict__ cell_start, const GridParams *__restrict__ gp,
                                                         const uint32_t *__restrict__ nbr_start, const uint32_t *__restrict__ nbr_count,
                                                         const uint16_t *__restrict__ list16, uint32_t *__restrict__ list32) {
    __shared__ TileRange rg[18];
    const GridParams g = *gp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int c = blockIdx.x; c < g.ncell; c += gridDim.x) {
        const uint32_t a0 = cell_start[c], a1 = cell_start[c + 1];
        const int c2 = c / (g.nc[0] * g.nc[1]);
        if (a0 == a1 || c2 < g.row_l0 || c2 >= g.row_l1) continue;  // block-uniform
        if (warp == 0) {
            TilePlan P;
            tile_plan(g, cell_start, c, a0, lane, P);
            if (lane < 9) { rg[2 * lane] = P.r0; rg[2 * lane + 1] = P.r1; }
        }
        __syncthreads();
        for (uint32_t i = a0 + (uint32_t)warp; i < a1; i += (uint32_t)n_warps) {
            const uint32_t s = nbr_start[i], cnt = nbr_count[i];
            for (uint32_t k = lane; k < cnt; k += 32) {
                const uint32_t t = list16[s + k];
                uint32_t slot = 0xffffffffu;
#pragma unroll 1
                for (int r = 0; r < 18; ++r)
                    if (t >= rg[r].off && t < rg[r].off + rg[r].cnt) { slot = rg[r].src + (t - rg[r].off); break; }
                list32[s + k] = slot;
            }
        }
        __syncthreads();
    }
}

}  // namespace

void launch_expand_rows(int grid_cells, const uint32_t *cell_start, const GridParams *g, const uint32_t *nbr_start,
                        const uint32_t *nbr_count, const uint16_t *list16, uint32_t *list32, cudaStream_t st, int64_t *launches) {
    const unsigned grid = (unsigned)std::max(1, std::min(grid_cells, 148 * 16));
    MC_LAUNCH(expand_rows_kernel, grid, 128, 0, st, cell_start, g, nbr_start, nbr_count, list16, list32);
    *launches += 1;
}

cudaError_t tile_sweep_prepare() {
    cudaError_t e = cudaSuccess;
#define MC_TB_ATTR(I, P) if (e == cudaSuccess) e = cudaFuncSetAttribute(tile_build_kernel<I, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    MC_TB_ATTR(uint32_t, false) MC_TB_ATTR(uint32_t, true) MC_TB_ATTR(uint16_t, false) MC_TB_ATTR(uint16_t, true)
#undef MC_TB_ATTR
#define MC_RB_ATTR(X) \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rows_build_kernel<2, TILE_WARPS, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rows_build_kernel<3, TILE_WARPS, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rows_build_kernel<1, RB_WARPS_DENSE, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RB_DENSE_BUDGET);
    MC_RB_ATTR(false) MC_RB_ATTR(true)
#undef MC_RB_ATTR
    return e;
}

// 16 B position + 4 B slot id per staged atom in 200 KB of shared memory; tiles above half of that run
// single-buffered (no copy/sweep overlap, but still one L2 read per cell instead of one per atom)
uint32_t tile_sweep_max_atoms() { return ((200u * 1024u) / 20u) & ~31u; }
// rows_build_kernel's one-CTA-per-SM configuration takes all the shared memory a block can have (227 KB less the kernel's
// few hundred static bytes): C3's largest 27-cell neighbourhood holds 11,355 atoms = 222 KB
uint32_t tile_sweep_max_atoms_dense() { return (uint32_t)(RB_DENSE_BUDGET / 20u) & ~31u; }

template <typename IDX, bool PART>
static void launch_tile_build_t(int n_rows, long long items, int split, int n_sms, const float4 *xyzq, const uint32_t *cell_start,
                                const GridParams *g, float rl2, float rc2_inner, const int *orig, const int32_t *excl_start,
                                const int32_t *excl_idx, uint32_t *nbr_count, uint32_t *nbr_start, IDX *nbr_list, uint32_t list_cap,
                                uint32_t tile_cap, uint32_t *ctl, cudaStream_t st) {
    // persistent: exactly one resident wave of CTAs pulls (cell, slice) items from ctl[0]
    const int n_stages = (size_t)2 * tile_cap * 20u <= 200u * 1024u ? 2 : 1;
    const size_t smem = (size_t)n_stages * tile_cap * (sizeof(float4) + sizeof(uint32_t));
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_build_kernel<IDX, PART>, (TILE_WARPS + 1) * 32, smem);
    if (per_sm < 1) per_sm = 1;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(items, (long long)n_sms * per_sm));
    cudaMemsetAsync(ctl, 0, 8 * sizeof(uint32_t), st);
    MC_LAUNCH(tile_build_kernel<IDX MC_COMMA PART>, grid, (TILE_WARPS + 1) * 32, smem, st, n_rows, xyzq, cell_start, g, rl2, rc2_inner, orig,
              excl_start, excl_idx, nbr_count, nbr_start, nbr_list, list_cap, tile_cap, split, n_stages, ctl);
}

// rows_build_kernel: tiles in flight and staged entries per row from what fits.  Three tiles in flight while the CTA stays
// small enough for three CTAs per SM, else two, else one; rows are staged when the longest row of the previous build (+ 25 %)
// fits next to the tiles, otherwise (first build of a system, very dense systems) every row takes the two-sweep path.
static void launch_rows_build(int n_rows, int grid_cells, long long items, int split, int n_sms, const float4 *xyzq, const uint32_t *cell_start,
                              const GridParams *g, float rl2, const int *orig, const int32_t *excl_start, const int32_t *excl_idx,
                              uint32_t *nbr_count, uint32_t *nbr_start, uint32_t *nbr_list, uint32_t list_cap, uint32_t tile_cap,
                              uint32_t row_hint, uint32_t *plan, uint32_t *ctl, cudaStream_t st, int min_blocks, const int *slot_of_orig) {
    const bool no_dense = min_blocks < 0;  // option rows_dense = 0 (A/B): the round-2 configuration, 8 consumer warps, two sweeps
    if (min_blocks < 0) min_blocks = -min_blocks;
    const size_t budget = 200u * 1024u, per_tile = (size_t)tile_cap * 20u;
    const uint32_t want_stage = row_hint ? ((row_hint + row_hint / 4u + 47u) & ~31u) : 0u;
    uint32_t stage_cap = want_stage, row_cap = 0u;
    size_t staging = (size_t)TILE_WARPS * RB_A * stage_cap * sizeof(uint16_t);
    if (per_tile + staging > budget) { stage_cap = 0u; staging = 0; }
    int n_stages = 1;
    const size_t per_cta = min_blocks == 2 ? 110u * 1024u : 72u * 1024u;  // shared memory that keeps min_blocks CTAs on an SM
    if (4 * per_tile + staging <= per_cta) n_stages = 4;
    else if (3 * per_tile + staging <= per_cta) n_stages = 3;
    else if (2 * per_tile + staging <= budget) n_stages = 2;
    // Tiles that leave room for ONE CTA per SM (dense systems: C3's 27 cells hold 7,400 atoms = 148 KB): twice the consumer
    // warps, and rows either staged (if that still fits) or written in a single direct sweep into row_cap entries each once
    // the previous build has told how long rows get (the first build of a system counts first and sweeps again).
    const bool dense = ((size_t)n_stages * per_tile + staging > 110u * 1024u || per_tile > budget) && !no_dense;
    if (dense) {
        const size_t staging16 = (size_t)RB_WARPS_DENSE * RB_A * want_stage * sizeof(uint16_t);
        if (want_stage && per_tile + staging16 <= RB_DENSE_BUDGET) { stage_cap = want_stage; staging = staging16; }
        else { stage_cap = 0u; staging = 0; row_cap = row_hint ? ((row_hint + row_hint / 4u + 15u) & ~7u) : 0u; }
        // list offsets are 32-bit: rows padded to row_cap must stay well inside them, else count first (exact rows)
        if ((unsigned long long)std::max(n_rows, 0) * row_cap > 3500000000ull) row_cap = 0u;
        n_stages = 2 * per_tile + staging <= RB_DENSE_BUDGET ? 2 : 1;
    }
    const size_t smem = (size_t)n_stages * per_tile + staging;
    cudaMemsetAsync(ctl, 0, 8 * sizeof(uint32_t), st);
    MC_LAUNCH(rows_plan_kernel, div_up((size_t)grid_cells * 32, 256), 256, 0, st, cell_start, g, plan, ctl);
#define MC_RB_GO(MINB_, NW_) do { if (excl_start) MC_RB_GO_X(MINB_, NW_, true); else MC_RB_GO_X(MINB_, NW_, false); } while (0)
#define MC_RB_GO_X(MINB_, NW_, X_) do { \
        int per_sm = 1; \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rows_build_kernel<MINB_, NW_, X_>, (NW_ + 1) * 32, smem); \
        if (per_sm < 1) per_sm = 1; \
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(items, (long long)n_sms * per_sm)); \
        MC_LAUNCH(rows_build_kernel<MINB_ MC_COMMA NW_ MC_COMMA X_>, grid, (NW_ + 1) * 32, smem, st, n_rows, xyzq, plan, g, rl2, orig, excl_start, excl_idx, \
                  nbr_count, nbr_start, nbr_list, list_cap, tile_cap, stage_cap, split, n_stages, ctl, row_cap, slot_of_orig); \
    } while (0)
    if (dense) MC_RB_GO(1, RB_WARPS_DENSE);
    else if (min_blocks == 2) MC_RB_GO(2, TILE_WARPS);
    else MC_RB_GO(3, TILE_WARPS);
#undef MC_RB_GO
#undef MC_RB_GO_X
}

size_t rows_plan_words(int grid_cells) { return (size_t)grid_cells * RB_PLAN_WORDS; }

void launch_tile_build(int n_rows, int grid_cells, int split, int n_sms, const float4 *xyzq, const uint32_t *cell_start,
                       const GridParams *g, float rl2, float rc2_inner, const int *orig, const int32_t *excl_start,
                       const int32_t *excl_idx, uint32_t *nbr_count, uint32_t *nbr_start, void *nbr_list, bool compact, bool partition,
                       uint32_t list_cap, uint32_t tile_cap, uint32_t *ctl, cudaStream_t st, int64_t *launches, int variant,
                       uint32_t row_hint, uint32_t *plan, int variant_min_blocks, const int *slot_of_orig) {
    const long long items = (long long)grid_cells * split;
    if (variant == 2 && !compact && !partition && plan) {
        launch_rows_build(n_rows, grid_cells, items, split, n_sms, xyzq, cell_start, g, rl2, orig, excl_start, excl_idx, nbr_count, nbr_start,
                          static_cast<uint32_t *>(nbr_list), list_cap, tile_cap, row_hint, plan, ctl, st, variant_min_blocks, slot_of_orig);
        *launches += 2;
        return;
    }
#define MC_TB_GO(I, P) launch_tile_build_t<I, P>(n_rows, items, split, n_sms, xyzq, cell_start, g, rl2, rc2_inner, orig, excl_start, excl_idx, \
                                                 nbr_count, nbr_start, static_cast<I *>(nbr_list), list_cap, tile_cap, ctl, st)
    if (compact) { if (partition) MC_TB_GO(uint16_t, true); else MC_TB_GO(uint16_t, false); }
    else { if (partition) MC_TB_GO(uint32_t, true); else MC_TB_GO(uint32_t, false); }
#undef MC_TB_GO
    *launches += 1;
}
