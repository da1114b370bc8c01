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
    const uint32_t *pw = lst32 + sub;
    uint32_t word = wi < nw ? *pw : 0u;
    if constexpr (COUL == MC_COULOMB_NONE && !ENERGY) {
        // packed path: two pairs per instruction (pair_terms.cuh)
        float2 c12 = c12_1, c6n = c6n_1;
        Acc2 b = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll 1
        while (wi < nw) {
            const uint32_t cur = word;
            const bool has1 = 2u * wi + 1u < cnt;
            wi += PT_LANES;
            pw += PT_LANES;
            if (wi < nw) word = *pw;
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
#pragma unroll 1
        while (wi < nw) {
            const uint32_t cur = word;
            const bool has1 = 2u * wi + 1u < cnt;
            wi += PT_LANES;
            pw += PT_LANES;
            if (wi < nw) word = *pw;
            const uint32_t j0 = cur & 0xffffu, j1 = has1 ? (cur >> 16) : j0;
            const float4 x0 = tile[j0], x1 = tile[j1];
            float2 l0 = lj1, l1 = lj1;
            if (MULTI) { l0 = row[ttype[j0]]; l1 = row[ttype[j1]]; }
            pair_term<COUL, WRAP, ENERGY>(xi, x0, l0, p, rc2_lj, a);
            if (has1) pair_term<COUL, WRAP, ENERGY>(xi, x1, l1, p, rc2_lj, a);
        }
    }
}

#define PT_REC_WORDS 64  // per cell: 18 range sources, 18 range lengths, then the scalars below
#define PT_REC_M 36
#define PT_REC_SELF 37
#define PT_REC_WRAP 38
#define PT_REC_A0 39
#define PT_REC_NA 40
#define PT_REC_BLK_SRC 41
#define PT_REC_BLK_ENTRIES 42
#define PT_REC_OK 43

// Everything the producer needs to stage a cell is fixed between two list builds: written once per build, one warp per
// cell, so that the force kernel's producer does ONE round trip to global memory per item (three independent coalesced
// loads) instead of a chain of five (work counter, cell bounds, stencil rows, row starts, row lengths).
//   plan[c]   : the tile layout of tile_ring.cuh (18 ranges in tile order), tile size, own offset, wrap class, the cell's
//               atoms, and its row block (one contiguous piece of the compact list: tile_build.cu allocates the rows of a
//               <= 32-atom cell with one cursor bump, in atom order)
//   rowtab[c] : per atom of the cell, (word offset of its row inside the block) << 16 | row length
__global__ void __launch_bounds__(128) cell_plan_kernel(const uint32_t *__restrict__ cell_start, const GridParams *__restrict__ gp,
                                                       const uint32_t *__restrict__ nbr_start, const uint32_t *__restrict__ nbr_count,
                                                       uint32_t tile_cap, uint32_t rows_cap, uint32_t *__restrict__ plan,
                                                       uint32_t *__restrict__ rowtab, uint32_t *__restrict__ ctl) {
    const GridParams g = *gp;
    const int lane = threadIdx.x & 31;
    const int c = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (c >= g.ncell) return;
    uint32_t *rec = plan + (size_t)c * PT_REC_WORDS;
    const uint32_t a0 = cell_start[c], a1 = cell_start[c + 1];
    const int c2 = c / (g.nc[0] * g.nc[1]);
    const uint32_t na = (c2 < g.row_l0 || c2 >= g.row_l1) ? 0u : a1 - a0;  // ghost layers carry no rows
    if (na == 0) {
        if (lane == 0) { rec[PT_REC_NA] = 0u; rec[PT_REC_OK] = 1u; }
        return;
    }
    TilePlan P;
    tile_plan(g, cell_start, c, a0, lane, P);
    if (lane < 9) {
        rec[2 * lane] = P.r0.src; rec[2 * lane + 1] = P.r1.src;
        rec[18 + 2 * lane] = P.r0.cnt; rec[18 + 2 * lane + 1] = P.r1.cnt;
    }
    uint32_t st = 0, cn = 0;
    if ((uint32_t)lane < na) { st = nbr_start[a0 + lane]; cn = nbr_count[a0 + lane]; }
    const uint32_t blk_src = __shfl_sync(MC_FULL_MASK, st, 0);
    const uint32_t end_l = st + ((cn + 7u) & ~7u);
    const uint32_t blk_entries = __shfl_sync(MC_FULL_MASK, end_l, (int)min(na, 32u) - 1) - blk_src;
    rowtab[(size_t)c * 32 + lane] = (((st - blk_src) >> 1) << 16) | (cn & 0xffffu);
    const bool okl = (uint32_t)lane >= na || (st >= blk_src && end_l - blk_src <= rows_cap && cn <= 0xffffu);
    const bool ok = __all_sync(MC_FULL_MASK, okl) && na <= 32u && blk_entries <= rows_cap && (blk_src & 7u) == 0u && P.m <= tile_cap;
    if (lane == 0) {
        rec[PT_REC_M] = P.m; rec[PT_REC_SELF] = P.self_off; rec[PT_REC_WRAP] = (uint32_t)P.wrap; rec[PT_REC_A0] = a0;
        rec[PT_REC_NA] = na; rec[PT_REC_BLK_SRC] = blk_src; rec[PT_REC_BLK_ENTRIES] = blk_entries; rec[PT_REC_OK] = ok ? 1u : 0u;
        if (!ok) ctl[3] = 2u;  // a list this kernel was not built for: the host falls back before any force launch
    }
}

struct PairTileArgs {
    const float4 *xyzq;
    const uint16_t *type;
    const GridParams *gp;
    const uint32_t *plan, *rowtab;   // cell_plan_kernel
    const uint16_t *list16;
    const float2 *ljtab;
    NbParams p;
    int lj_on;
    float4 *force;
    uint32_t tile_cap;   // atoms per stage (multiple of 32)
    uint32_t rows_cap;   // 16-bit entries of the row block a stage holds (multiple of 8)
    int n_stages;
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

    // item order: layers that need no ghost first, then the first and the last row layer (decomposed ranks only);
    // items are dealt round-robin over the CTAs (the cost of a cell varies little: static scheduling, no work counter)
    const int plane = g.nc[0] * g.nc[1];
    const int nl = g.row_l1 - g.row_l0;
    const bool halo = A.wait.ready_prev != nullptr;
    const long long n_int = halo ? (long long)max(nl - 2, 0) * plane : (long long)nl * plane;
    const long long n_items = halo ? n_int + (long long)min(nl, 2) * plane : n_int;

    if (warp == 0) {
        // ===== producer =====
        bool waited = !halo;
        int p_stage = 0;
        uint32_t p_phase = 0u;
        auto cell_of = [&](long long w, bool &boundary) -> int {
            boundary = false;
            if (!halo) return g.row_l0 * plane + (int)w;
            if (w < n_int) return (g.row_l0 + 1) * plane + (int)w;
            const long long w2 = w - n_int;
            boundary = true;
            return w2 < plane ? g.row_l0 * plane + (int)w2 : (g.row_l1 - 1) * plane + (int)(w2 - plane);
        };
        // the record of the NEXT item is requested before the current one is staged: the one global round trip of the
        // producer hides behind the wait for a free stage
        long long w = blockIdx.x;
        uint32_t rec_lo = 0, rec_hi = 0, tab = 0;
        bool boundary = false;
        if (w < n_items) {
            const int c = cell_of(w, boundary);
            rec_lo = __ldg(A.plan + (size_t)c * PT_REC_WORDS + lane);
            rec_hi = __ldg(A.plan + (size_t)c * PT_REC_WORDS + 32 + lane);
            tab = __ldg(A.rowtab + (size_t)c * 32 + lane);
        }
        for (;;) {
            const bool done = w >= n_items;
            const uint32_t cur_lo = rec_lo, cur_hi = rec_hi, cur_tab = tab;
            const bool cur_boundary = boundary;
            const long long wn = w + gridDim.x;
            if (!done && wn < n_items) {
                const int cn = cell_of(wn, boundary);
                rec_lo = __ldg(A.plan + (size_t)cn * PT_REC_WORDS + lane);
                rec_hi = __ldg(A.plan + (size_t)cn * PT_REC_WORDS + 32 + lane);
                tab = __ldg(A.rowtab + (size_t)cn * 32 + lane);
            }
            w = wn;
            // lanes 0..17 own the 18 ranges: source in word l, length in word 18 + l (= lanes 18..31 of rec_lo, 0..3 of rec_hi)
            const uint32_t src = cur_lo;
            const uint32_t cnt_a = __shfl_sync(MC_FULL_MASK, cur_lo, (lane + 18) & 31), cnt_b = __shfl_sync(MC_FULL_MASK, cur_hi, (lane + 18) & 31);
            uint32_t cnt = lane < 14 ? cnt_a : cnt_b;
            const uint32_t m = __shfl_sync(MC_FULL_MASK, cur_hi, PT_REC_M - 32), self_off = __shfl_sync(MC_FULL_MASK, cur_hi, PT_REC_SELF - 32),
                           wrap = __shfl_sync(MC_FULL_MASK, cur_hi, PT_REC_WRAP - 32), a0 = __shfl_sync(MC_FULL_MASK, cur_hi, PT_REC_A0 - 32),
                           na = __shfl_sync(MC_FULL_MASK, cur_hi, PT_REC_NA - 32), blk_src = __shfl_sync(MC_FULL_MASK, cur_hi, PT_REC_BLK_SRC - 32),
                           blk_entries = __shfl_sync(MC_FULL_MASK, cur_hi, PT_REC_BLK_ENTRIES - 32), ok = __shfl_sync(MC_FULL_MASK, cur_hi, PT_REC_OK - 32);
            if (!done && (na == 0u || !ok)) continue;  // empty cell / ghost layer (warp-uniform); !ok was reported by cell_plan_kernel
            if (done || lane >= 18) cnt = 0u;
            uint32_t inc = cnt;  // exclusive prefix of the range lengths = tile offsets
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(MC_FULL_MASK, inc, d);
                if (lane >= d) inc += v;
            }
            const uint32_t off = inc - cnt;
            if (!done && cur_boundary && !waited) {
                if (lane == 0) {
                    halo_spin(A.wait.ready_prev, A.wait.want, A.wait.err);
                    halo_spin(A.wait.ready_next, A.wait.want, A.wait.err);
                    fence_proxy_async();  // the peers' stores (acquired above) before the bulk-copy reads
                }
                __syncwarp();
                waited = true;
            }
            const int s = p_stage;
            mbar_wait_parked(&empty_bar[s], p_phase ^ 1u, 20000u);
            if (++p_stage == n_stages) { p_stage = 0; p_phase ^= 1u; }
            unsigned char *stage = stage0 + (size_t)s * stage_bytes;
            float4 *tile = reinterpret_cast<float4 *>(stage);
            if (lane == 0) {
                meta[s].m = done ? 0u : m; meta[s].a0 = done ? 0xffffffffu : a0; meta[s].a1 = done ? 0xffffffffu : a0 + na;
                meta[s].self_off = self_off; meta[s].wrap = (int)wrap;
            }
            row_tab[s][lane] = cur_tab;
            if (MULTI && !done) {
                uint16_t *ttype = reinterpret_cast<uint16_t *>(stage + type_off);
                for (int src_lane = 0; src_lane < 18; ++src_lane) {
                    const uint32_t s0 = __shfl_sync(MC_FULL_MASK, src, src_lane), n0 = __shfl_sync(MC_FULL_MASK, cnt, src_lane),
                                   o0 = __shfl_sync(MC_FULL_MASK, off, src_lane);
                    for (uint32_t t = lane; t < n0; t += 32) ttype[o0 + t] = A.type[s0 + t];
                }
            }
            __syncwarp();
#ifdef MC_PT_EXPERIMENT_ONE_COPY  // (timing experiment: one tile copy per item instead of ~9; the data is wrong on purpose)
            const uint32_t cnt0 = __shfl_sync(MC_FULL_MASK, cnt, 0);
            if (lane == 0) mbar_expect_tx(&full_bar[s], done ? 0u : cnt0 * (uint32_t)sizeof(float4) + blk_entries * (uint32_t)sizeof(uint16_t));
            if (lane != 0) cnt = 0u;
#else
            if (lane == 0)  // release: meta, row table (+ types) visible
                mbar_expect_tx(&full_bar[s], done ? 0u : m * (uint32_t)sizeof(float4) + blk_entries * (uint32_t)sizeof(uint16_t));
#endif
            __syncwarp();
            if (!done) {
                if (cnt) tma_bulk_g2s(tile + off, A.xyzq + src, cnt * (uint32_t)sizeof(float4), &full_bar[s]);
                if (lane == 18 && blk_entries)
                    tma_bulk_g2s(stage + rows_off, A.list16 + blk_src, blk_entries * (uint32_t)sizeof(uint16_t), &full_bar[s]);
            }
            if (done) break;
        }
    } else {
        // ===== consumers =====
        const int cw = warp - 1;
        const int sub = lane % PT_LANES, rsub = lane / PT_LANES;
        const float rc2_lj = A.lj_on ? A.p.rc2_lj : -1.f;
        const float s6c = A.p.sig2 * A.p.sig2 * A.p.sig2;  // single-type constants of the packed path
        const float2 c12_1 = make_float2(2.f * A.p.eps24 * s6c * s6c, 2.f * A.p.eps24 * s6c * s6c), c6n_1 = make_float2(-A.p.eps24 * s6c, -A.p.eps24 * s6c);
        int s = 0;
        uint32_t phase = 0u, rot = (uint32_t)cw;  // rot: this warp's first quad of the current item (rotates with the items)
        for (;; ) {
            mbar_wait_parked(&full_bar[s], phase, 20000u);
            const PtMeta M = meta[s];
            if (M.a0 == 0xffffffffu) break;
            const unsigned char *stage = stage0 + (size_t)s * stage_bytes;
            const float4 *tile = reinterpret_cast<const float4 *>(stage);
            const uint32_t *rows_s = reinterpret_cast<const uint32_t *>(stage + rows_off);
            const uint16_t *ttype = reinterpret_cast<const uint16_t *>(stage + type_off);
            const uint32_t na = M.a1 - M.a0;
            const uint32_t nq = (na + PT_RPW - 1) / PT_RPW;
            // row quads are dealt round-robin, rotated by the item number: a ~19-atom cell has 5 quads for the consumer warps
#pragma unroll 1
            for (uint32_t q = rot; q < nq; q += PT_WARPS) {
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
#ifndef MC_PT_EXPERIMENT_NO_MATH  // (timing experiment: what the staging pipeline alone costs)
                if (M.wrap) row_loop_tile<MULTI, COUL, true, ENERGY>(xi, lst32, cnt, sub, tile, ttype, row, A.p, rc2_lj, c12_1, c6n_1, a);
                else row_loop_tile<MULTI, COUL, false, ENERGY>(xi, lst32, cnt, sub, tile, ttype, row, A.p, rc2_lj, c12_1, c6n_1, a);
#else
                a.fx = xi.x + (float)cnt + (float)lst32[sub];
#endif
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
            if (++s == n_stages) { s = 0; phase ^= 1u; }
            if (++rot == PT_WARPS) rot = 0u;
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
size_t pair_tile_smem(uint32_t tile_cap, uint32_t rows_max_entries, int n_types, bool multi, int *n_stages_out, uint32_t *rows_cap_out,
                      int force_stages) {
    const size_t tab = multi ? (((size_t)n_types * n_types * sizeof(float2) + 127) & ~(size_t)127) : 0;
    const size_t tile_b = (size_t)tile_cap * (sizeof(float4) + (multi ? sizeof(uint16_t) : 0));
    const uint32_t rows_cap = (rows_max_entries + 7u) & ~7u;
    const size_t stage = (tile_b + (size_t)rows_cap * 2 + 15) & ~(size_t)15;
    if (stage > 40u * 1024u || tab + 2 * stage > 200u * 1024u) return 0;
    // as many stages in flight as keep ~44 KB per CTA (five CTAs per SM), at least two
#ifdef MC_PT_EXPERIMENT_STAGES
    int ns = MC_PT_EXPERIMENT_STAGES;
#else
    int ns = stage * 4 + tab <= 44u * 1024u ? 4 : (stage * 3 + tab <= 44u * 1024u ? 3 : 2);
#endif
    // option pair_tile_stages: measured on C4 (profiles/pair_tile_r2_experiments.txt) three tiles in flight at four CTAs per SM
    // beat both two tiles at five CTAs and four tiles at three
    if (force_stages >= 2 && force_stages <= PT_MAX_STAGES && tab + (size_t)force_stages * stage <= 200u * 1024u) ns = force_stages;
    if (n_stages_out) *n_stages_out = ns;
    if (rows_cap_out) *rows_cap_out = rows_cap;
    return tab + (size_t)ns * stage;
}

void launch_cell_plan(int n_cells, const uint32_t *cell_start, const GridParams *g, const uint32_t *nbr_start, const uint32_t *nbr_count,
                      uint32_t tile_cap, uint32_t rows_cap, uint32_t *plan, uint32_t *rowtab, uint32_t *ctl, cudaStream_t st, int64_t *launches) {
    if (n_cells <= 0) return;
    MC_LAUNCH(cell_plan_kernel, div_up((size_t)n_cells * 32, 128), 128, 0, st, cell_start, g, nbr_start, nbr_count, tile_cap, rows_cap, plan,
              rowtab, ctl);
    *launches += 1;
}

size_t pair_tile_plan_words() { return PT_REC_WORDS; }

void launch_pair_tile(const PairTileLaunch &L, cudaStream_t st, int64_t *launches) {
    PairTileArgs A;
    A.xyzq = L.xyzq; A.type = L.type; A.gp = L.grid; A.plan = L.plan; A.rowtab = L.rowtab;
    A.list16 = L.list16; A.ljtab = L.ljtab;
    A.p = L.p; A.lj_on = L.lj_on; A.force = L.force; A.tile_cap = L.tile_cap; A.wait = L.wait;
    int ns = 1;
    uint32_t rows_cap = 0;
    const size_t smem = pair_tile_smem(L.tile_cap, L.rows_max_entries, L.p.n_types, L.multi, &ns, &rows_cap, L.force_stages);
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
