// pair_tile.cu -- the nonbonded pair-force inner loop with the neighbour tile staged in shared memory by TMA
// (BASELINE north_star: "neighbour tiles staged into shared memory via TMA, warp-shuffle partial-force
// reductions").  Same arithmetic as pair_force.cu (pair_terms.cuh; reference src/cuda/util.cu:54-139, cuda.cu:73-102),
// different data movement:
//
//   pair_force.cu   rows of 32-bit GLOBAL slots, one 16-byte __ldg gather per listed pair: measured L1TEX-data-pipe
//                   bound (92.7 %, ~13.6 wavefronts per warp-wide gather = distinct 128-byte lines under 32 addresses)
//   pair_tile.cu    persistent, warp-specialised.  Work item = one cell.  A producer warp brings the cell's 27-cell
//                   candidate tile (~500 atoms, <= 18 contiguous ranges of the cell-ordered xyzq array, the SAME layout
//                   the list build used: tile_ring.cuh) into shared memory with cp.async.bulk (SASS UBLKCP) on an
//                   mbarrier ring; consumer warps run the cell's rows, whose entries are 16-bit TILE-LOCAL indices
//                   (half the index stream of HBM: 2 B instead of 4 B per listed pair), and gather xyzq_j with LDS.128
//                   from the tile -- no L1 lines, no sector waste; the own atom comes from the tile as well.
//
// Decomposed ranks: items of the interior layers are handed out first; before the first tile of a boundary layer (it
// contains ghost atoms the neighbours' kick_drift kernels store over NVLink) the producer waits for this epoch's ready
// flags (halo_sync.cuh) -- one thread per CTA, no L1 to invalidate: the bulk copy reads through L2.
//
// Roofline: HBM by SURVEY 8d's count (32 N + 20 P_full algorithmic bytes); physically the kernel is bound by the
// shared-memory / L1 data pipe and instruction issue, DRAM carries 2 B per listed pair + one pass over xyzq.
#include <algorithm>

#include "common.cuh"
#include "pair_force.cuh"
#include "pair_terms.cuh"
#include "tile_ring.cuh"

namespace {

#ifndef MC_PT_WARPS
#define MC_PT_WARPS 7  // consumer warps per CTA (7 + the producer = 256 threads: five CTAs per SM at 48 registers)
#endif
#ifndef MC_PT_MIN_BLOCKS
#define MC_PT_MIN_BLOCKS 5
#endif
constexpr int PT_WARPS = MC_PT_WARPS;
constexpr int PT_MAX_STAGES = 4;
constexpr int PT_LANES = 8;             // lanes per row
constexpr int PT_RPW = 32 / PT_LANES;   // rows per warp pass

struct PtMeta {
    uint32_t m, a0, a1, self_off;  // a0 == 0xffffffff: no more work
    int wrap;
};

// One row: entries k = 2 sub, 2 sub + 1 of every group of 2 LANES entries -- a lane reads its two 16-bit indices as one
// 32-bit word (8 lanes = 32 consecutive bytes of the index stream), gathers both atoms from the tile, then does the math.
// The row itself sits in shared memory as well (staged by the producer with the tile): the loop touches no global memory.  An odd row length leaves the second entry of the last word
// as padding: it is gathered from the first entry's slot and masked by a cutoff below zero.
template <bool MULTI, int COUL, bool WRAP, bool ENERGY>
__device__ __forceinline__ void row_loop_tile(const float4 xi, const uint32_t *lst32, uint32_t cnt, int sub,
                                              const float4 *tile, const uint16_t *ttype, const float2 *row,
                                              const NbParams &p, const float rc2_lj, const float2 c12_1, const float2 c6n_1, Acc &a) {
    const uint32_t nw = (cnt + 1u) >> 1;  // index words of this row
    uint32_t wi = (uint32_t)sub;
    uint32_t word = wi < nw ? lst32[wi] : 0u;
    if constexpr (COUL == MC_COULOMB_NONE && !ENERGY) {
        // packed path: two pairs per instruction (pair_terms.cuh)
        float2 c12 = c12_1, c6n = c6n_1;
        Acc2 b = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        while (wi < nw) {
            const uint32_t cur = word;
            const bool has1 = 2u * wi + 1u < cnt;
            wi += PT_LANES;
            if (wi < nw) word = lst32[wi];
            const uint32_t j0 = cur & 0xffffu, j1 = has1 ? (cur >> 16) : j0;
            const float4 x0 = tile[j0], x1 = tile[j1];
            if (MULTI) {
                const float2 l0 = row[ttype[j0]], l1 = row[ttype[j1]];  // (c12, -c6) per type pair (staged by the kernel)
                c12 = make_float2(l0.x, l1.x);
                c6n = make_float2(l0.y, l1.y);
            }
            pair_term2_lj<WRAP>(xi, x0, x1, c12, c6n, p, rc2_lj, has1 ? rc2_lj : -1.f, b);
        }
        // d was x_j - x_i: flip the sign once
        a.fx -= b.fx.x + b.fx.y;
        a.fy -= b.fy.x + b.fy.y;
        a.fz -= b.fz.x + b.fz.y;
    } else {
        const float2 lj1 = make_float2(p.sig2, p.eps24);
        while (wi < nw) {
            const uint32_t cur = word;
            const bool has1 = 2u * wi + 1u < cnt;
            wi += PT_LANES;
            if (wi < nw) word = lst32[wi];
            const uint32_t j0 = cur & 0xffffu, j1 = has1 ? (cur >> 16) : j0;
            const float4 x0 = tile[j0], x1 = tile[j1];
            float2 l0 = lj1, l1 = lj1;
            if (MULTI) { l0 = row[ttype[j0]]; l1 = row[ttype[j1]]; }
            pair_term<COUL, WRAP, ENERGY>(xi, x0, l0, p, rc2_lj, a);
            if (has1) pair_term<COUL, WRAP, ENERGY>(xi, x1, l1, p, rc2_lj, a);
        }
    }
}

struct PairTileArgs {
    const float4 *xyzq;
    const uint16_t *type;
    const uint32_t *cell_start;
    const GridParams *gp;
    const uint32_t *nbr_start, *nbr_count;
    const uint16_t *list16;
    const float2 *ljtab;
    NbParams p;
    int lj_on;
    float4 *force;
    uint32_t tile_cap;   // atoms per stage (multiple of 32)
    uint32_t rows_cap;   // 16-bit entries of the row block a stage holds (multiple of 8)
    int n_stages;
    uint32_t *ctl;       // [0] work counter, [1] CTAs that have drained, [3] a tile did not fit (never, if the build fitted)
    HaloWait wait;       // decomposed rank, fused halo: ready flags of this epoch (ready_prev == nullptr: none)
};

template <bool MULTI, int COUL, bool ENERGY>
__global__ void __launch_bounds__((PT_WARPS + 1) * 32, MC_PT_MIN_BLOCKS) pair_tile_kernel(const PairTileArgs A) {
    MC_DYN_SHARED_ALIGNED(unsigned char, smem_raw, 128);
    __shared__ __align__(8) uint64_t full_bar[PT_MAX_STAGES], empty_bar[PT_MAX_STAGES];
    __shared__ PtMeta meta[PT_MAX_STAGES];
    __shared__ uint32_t row_tab[PT_MAX_STAGES][32];  // per staged row: word offset inside the block << 16 | entries
    // dynamic shared memory: [LJ table (MULTI)] then per stage: tile_cap float4 positions, rows_cap 16-bit list entries
    // (the cell's row block), tile_cap u16 types when MULTI
    const int nt2 = MULTI ? A.p.n_types * A.p.n_types : 0;
    float2 *s_tab = reinterpret_cast<float2 *>(smem_raw);
    unsigned char *stage0 = smem_raw + (((size_t)nt2 * sizeof(float2) + 127) & ~(size_t)127);
    const size_t rows_off = (size_t)A.tile_cap * sizeof(float4);
    const size_t type_off = rows_off + (size_t)A.rows_cap * sizeof(uint16_t);
    const size_t stage_bytes = (type_off + (MULTI ? (size_t)A.tile_cap * sizeof(uint16_t) : 0) + 15) & ~(size_t)15;
    const int n_stages = A.n_stages;

    const GridParams g = *A.gp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], PT_WARPS);
        }
    }
    if (MULTI)
        for (int t = threadIdx.x; t < nt2; t += blockDim.x) {
            float2 l = A.ljtab[t];  // (sigma^2, 24 eps)
            if (COUL == MC_COULOMB_NONE && !ENERGY) {  // the packed LJ path takes (48 eps sigma^12, -24 eps sigma^6)
                const float s6c = l.x * l.x * l.x;
                l = make_float2(2.f * l.y * s6c * s6c, -l.y * s6c);
            }
            s_tab[t] = l;
        }
    __syncthreads();

    // item order: layers that need no ghost first, then the first and the last row layer (decomposed ranks only)
    const int plane = g.nc[0] * g.nc[1];
    const int nl = g.row_l1 - g.row_l0;
    const bool halo = A.wait.ready_prev != nullptr;
    const long long n_int = halo ? (long long)max(nl - 2, 0) * plane : (long long)nl * plane;
    const long long n_items = halo ? n_int + (long long)min(nl, 2) * plane : n_int;

    if (warp == 0) {
        // ===== producer =====
        bool waited = !halo;
        uint32_t it = 0;
        for (;;) {
            long long w = 0;
            if (lane == 0) w = (long long)atomicAdd(A.ctl, 1u);
            w = __shfl_sync(MC_FULL_MASK, w, 0);
            const bool done = w >= n_items;
            uint32_t a0 = 0xffffffffu, a1 = 0xffffffffu;
            TilePlan P;
            P.m = 0; P.self_off = 0; P.wrap = 0; P.r0 = P.r1 = TileRange{0u, 0u, 0u};
            uint32_t blk_src = 0, blk_entries = 0, my_tab = 0;
            if (!done) {
                int c;
                bool boundary = false;
                if (!halo) c = g.row_l0 * plane + (int)w;
                else if (w < n_int) c = (g.row_l0 + 1) * plane + (int)w;
                else {
                    const long long w2 = w - n_int;
                    c = w2 < plane ? g.row_l0 * plane + (int)w2 : (g.row_l1 - 1) * plane + (int)(w2 - plane);
                    boundary = true;
                }
                a0 = A.cell_start[c];
                a1 = A.cell_start[c + 1];
                if (a0 == a1) continue;  // empty cell (warp-uniform)
                tile_plan(g, A.cell_start, c, a0, lane, P);
                if (P.m > A.tile_cap) {  // cannot happen when the list was built with this layout; never read out of bounds
                    if (lane == 0) A.ctl[3] = 1u;
                    continue;
                }
                // The rows of a cell (<= 32 atoms: the launcher checked the build's maxima) are ONE contiguous block of the list
                // -- tile_build.cu allocates them with one cursor bump, in atom order.  The block travels with the tile, and
                // with it an (offset, count) table, so that the consumers never wait for global memory.
                const uint32_t na = a1 - a0;
                uint32_t st = 0, cn = 0;
                if ((uint32_t)lane < na) { st = __ldg(A.nbr_start + a0 + lane); cn = __ldg(A.nbr_count + a0 + lane); }
                blk_src = __shfl_sync(MC_FULL_MASK, st, 0);
                const uint32_t end_l = st + ((cn + 7u) & ~7u);
                blk_entries = __shfl_sync(MC_FULL_MASK, end_l, (int)min(na, 32u) - 1) - blk_src;
                my_tab = (((st - blk_src) >> 1) << 16) | (cn & 0xffffu);
                const bool okl = (uint32_t)lane >= na || (st >= blk_src && end_l - blk_src <= A.rows_cap && cn <= 0xffffu);
                if (!__all_sync(MC_FULL_MASK, okl) || na > 32u || blk_entries > A.rows_cap || (blk_src & 7u) != 0u) {
                    if (lane == 0) A.ctl[3] = 2u;  // a list this kernel was not built for: reported, never read out of bounds
                    continue;
                }
                if (boundary && !waited) {
                    if (lane == 0) {
                        halo_spin(A.wait.ready_prev, A.wait.want, A.wait.err);
                        halo_spin(A.wait.ready_next, A.wait.want, A.wait.err);
                        fence_proxy_async();  // the peers' stores (acquired above) before this thread's bulk-copy reads
                    }
                    __syncwarp();
                    waited = true;
                }
            }
            const int s = (int)(it % (uint32_t)n_stages);
            mbar_wait(&empty_bar[s], ((it / (uint32_t)n_stages) & 1u) ^ 1u);
            unsigned char *stage = stage0 + (size_t)s * stage_bytes;
            float4 *tile = reinterpret_cast<float4 *>(stage);
            if (lane == 0) {
                meta[s].m = P.m; meta[s].a0 = a0; meta[s].a1 = a1; meta[s].self_off = P.self_off; meta[s].wrap = P.wrap;
            }
            row_tab[s][lane] = my_tab;
            if (MULTI && !done) {
                uint16_t *ttype = reinterpret_cast<uint16_t *>(stage + type_off);
                for (int src_lane = 0; src_lane < 9; ++src_lane) {
                    const uint32_t s0 = __shfl_sync(MC_FULL_MASK, P.r0.src, src_lane), n0 = __shfl_sync(MC_FULL_MASK, P.r0.cnt, src_lane),
                                   o0 = __shfl_sync(MC_FULL_MASK, P.r0.off, src_lane), s1 = __shfl_sync(MC_FULL_MASK, P.r1.src, src_lane),
                                   n1 = __shfl_sync(MC_FULL_MASK, P.r1.cnt, src_lane), o1 = __shfl_sync(MC_FULL_MASK, P.r1.off, src_lane);
                    for (uint32_t t = lane; t < n0; t += 32) ttype[o0 + t] = A.type[s0 + t];
                    for (uint32_t t = lane; t < n1; t += 32) ttype[o1 + t] = A.type[s1 + t];
                }
            }
            __syncwarp();
            if (lane == 0)  // release: meta, row table (+ types) visible
                mbar_expect_tx(&full_bar[s], P.m * (uint32_t)sizeof(float4) + blk_entries * (uint32_t)sizeof(uint16_t));
            __syncwarp();
            if (lane < 9) {
                if (P.r0.cnt) tma_bulk_g2s(tile + P.r0.off, A.xyzq + P.r0.src, P.r0.cnt * (uint32_t)sizeof(float4), &full_bar[s]);
                if (P.r1.cnt) tma_bulk_g2s(tile + P.r1.off, A.xyzq + P.r1.src, P.r1.cnt * (uint32_t)sizeof(float4), &full_bar[s]);
            }
            if (lane == 9 && blk_entries)
                tma_bulk_g2s(stage + rows_off, A.list16 + blk_src, blk_entries * (uint32_t)sizeof(uint16_t), &full_bar[s]);
            ++it;
            if (done) break;
        }
    } else {
        // ===== consumers =====
        const int cw = warp - 1;
        const int sub = lane % PT_LANES, rsub = lane / PT_LANES;
        const float rc2_lj = A.lj_on ? A.p.rc2_lj : -1.f;
        const float s6c = A.p.sig2 * A.p.sig2 * A.p.sig2;  // single-type constants of the packed path
        const float2 c12_1 = make_float2(2.f * A.p.eps24 * s6c * s6c, 2.f * A.p.eps24 * s6c * s6c), c6n_1 = make_float2(-A.p.eps24 * s6c, -A.p.eps24 * s6c);
        for (uint32_t it = 0;; ++it) {
            const int s = (int)(it % (uint32_t)n_stages);
            mbar_wait(&full_bar[s], (it / (uint32_t)n_stages) & 1u);
            const PtMeta M = meta[s];
            if (M.a0 == 0xffffffffu) break;
            const unsigned char *stage = stage0 + (size_t)s * stage_bytes;
            const float4 *tile = reinterpret_cast<const float4 *>(stage);
            const uint32_t *rows_s = reinterpret_cast<const uint32_t *>(stage + rows_off);
            const uint16_t *ttype = reinterpret_cast<const uint16_t *>(stage + type_off);
            const uint32_t na = M.a1 - M.a0;
            const uint32_t nq = (na + PT_RPW - 1) / PT_RPW;
            // row quads are dealt round-robin, rotated by the item number: a ~19-atom cell has 5 quads for 8 warps
            for (uint32_t q = (uint32_t)(cw + (int)(it % PT_WARPS)) % PT_WARPS; q < nq; q += PT_WARPS) {
                const uint32_t r = q * PT_RPW + (uint32_t)rsub;
                const bool live = r < na;
                const uint32_t rr = live ? r : 0u;
                const uint32_t i = M.a0 + rr;
                Acc a = {0.f, 0.f, 0.f, 0.f};
                const float4 xi = tile[M.self_off + rr];
                const float2 *row = MULTI ? s_tab + (int)ttype[M.self_off + rr] * A.p.n_types : nullptr;
                const uint32_t tab = row_tab[s][rr];
                const uint32_t cnt = live ? (tab & 0xffffu) : 0u;
                const uint32_t *lst32 = rows_s + (tab >> 16);
                if (M.wrap) row_loop_tile<MULTI, COUL, true, ENERGY>(xi, lst32, cnt, sub, tile, ttype, row, A.p, rc2_lj, c12_1, c6n_1, a);
                else row_loop_tile<MULTI, COUL, false, ENERGY>(xi, lst32, cnt, sub, tile, ttype, row, A.p, rc2_lj, c12_1, c6n_1, a);
                // warp-shuffle partial-force reduction across the lanes of this row
#pragma unroll
                for (int d = PT_LANES / 2; d > 0; d >>= 1) {
                    a.fx += __shfl_xor_sync(MC_FULL_MASK, a.fx, d);
                    a.fy += __shfl_xor_sync(MC_FULL_MASK, a.fy, d);
                    a.fz += __shfl_xor_sync(MC_FULL_MASK, a.fz, d);
                    if (ENERGY) a.e += __shfl_xor_sync(MC_FULL_MASK, a.e, d);
                }
                if (live && sub == 0) A.force[i] = make_float4(a.fx, a.fy, a.fz, a.e);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
        // The last CTA to drain re-arms the work counter for the next launch (no memset between the step kernels).  A
        // CTA has drained when its consumers have seen the end marker: every atomicAdd on ctl[0] is behind it.
        if (cw == 0 && lane == 0) {
            __threadfence();
            if (atomicAdd(A.ctl + 1, 1u) == gridDim.x - 1) {
                A.ctl[0] = 0u;
                A.ctl[1] = 0u;
                __threadfence();
            }
        }
    }
}

template <bool MULTI, int COUL, bool ENERGY>
void launch_one(const PairTileLaunch &L, const PairTileArgs &A, size_t smem, cudaStream_t st) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pair_tile_kernel<MULTI, COUL, ENERGY>, (PT_WARPS + 1) * 32, smem);
    if (per_sm < 1) per_sm = 1;
    const long long items = std::max(1, L.grid_cells);
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(items, (long long)L.n_sms * per_sm));
    MC_LAUNCH(pair_tile_kernel<MULTI MC_COMMA COUL MC_COMMA ENERGY>, grid, (PT_WARPS + 1) * 32, smem, st, A);
}

}  // namespace

#ifdef MC_HAVE_LAUNCH
cudaError_t pair_tile_prepare() {
    cudaError_t e = cudaSuccess;
#define MC_ATTR(M, C, E) \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pair_tile_kernel<M, C, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
#define MC_ATTR_C(M) MC_ATTR(M, MC_COULOMB_NONE, true) MC_ATTR(M, MC_COULOMB_NONE, false) MC_ATTR(M, MC_COULOMB_PLAIN, true) \
    MC_ATTR(M, MC_COULOMB_PLAIN, false) MC_ATTR(M, MC_COULOMB_ERFC, true) MC_ATTR(M, MC_COULOMB_ERFC, false)
    MC_ATTR_C(true) MC_ATTR_C(false)
#undef MC_ATTR_C
#undef MC_ATTR
    return e;
}

// Shared memory the kernel needs; 0 = this system is not for it (use pair_force.cu): the kernel stages, per cell, the
// 27-cell tile AND the cell's rows, which must be one block of <= 32 rows (max_cell_atoms) -- LJ-fluid-like systems, ~12 KB
// of tile + ~5 KB of rows.  Dense / long-cutoff systems (hundreds of atoms per cell, rows of a thousand entries) are not.
size_t pair_tile_smem(uint32_t tile_cap, uint32_t rows_max_entries, int n_types, bool multi, int *n_stages_out, uint32_t *rows_cap_out) {
    const size_t tab = multi ? (((size_t)n_types * n_types * sizeof(float2) + 127) & ~(size_t)127) : 0;
    const size_t tile_b = (size_t)tile_cap * (sizeof(float4) + (multi ? sizeof(uint16_t) : 0));
    const uint32_t rows_cap = (rows_max_entries + 7u) & ~7u;
    const size_t stage = (tile_b + (size_t)rows_cap * 2 + 15) & ~(size_t)15;
    if (stage > 40u * 1024u || tab + 2 * stage > 200u * 1024u) return 0;
    // as many stages in flight as keep ~40 KB per CTA (five CTAs per SM), at least two
    const int ns = stage * 3 + tab <= 40u * 1024u ? 3 : 2;
    if (n_stages_out) *n_stages_out = ns;
    if (rows_cap_out) *rows_cap_out = rows_cap;
    return tab + (size_t)ns * stage;
}

void launch_pair_tile(const PairTileLaunch &L, cudaStream_t st, int64_t *launches) {
    PairTileArgs A;
    A.xyzq = L.xyzq; A.type = L.type; A.cell_start = L.cell_start; A.gp = L.grid;
    A.nbr_start = L.nbr_start; A.nbr_count = L.nbr_count; A.list16 = L.list16; A.ljtab = L.ljtab;
    A.p = L.p; A.lj_on = L.lj_on; A.force = L.force; A.tile_cap = L.tile_cap; A.ctl = L.ctl; A.wait = L.wait;
    int ns = 1;
    uint32_t rows_cap = 0;
    const size_t smem = pair_tile_smem(L.tile_cap, L.rows_max_entries, L.p.n_types, L.multi, &ns, &rows_cap);
    A.n_stages = ns;
    A.rows_cap = rows_cap;
#define MC_PT_E(M, C) \
    if (L.energy) launch_one<M, C, true>(L, A, smem, st); else launch_one<M, C, false>(L, A, smem, st)
#define MC_PT_C(M)                                                  \
    switch (L.coul) {                                               \
        case MC_COULOMB_NONE: MC_PT_E(M, MC_COULOMB_NONE); break;   \
        case MC_COULOMB_PLAIN: MC_PT_E(M, MC_COULOMB_PLAIN); break; \
        default: MC_PT_E(M, MC_COULOMB_ERFC); break;                \
    }
    if (L.multi) { MC_PT_C(true) } else { MC_PT_C(false) }
#undef MC_PT_C
#undef MC_PT_E
    *launches += 1;
}
#endif
