// tile_ring.cuh -- what the two TMA-staged, warp-specialised kernels share (tile_build.cu: Verlet-list build,
// pair_tile.cu: pair forces): the mbarrier / bulk-copy wrappers and the LAYOUT OF A CELL'S CANDIDATE TILE.
//
// The tile of cell c is the concatenation of the <= 18 contiguous slot ranges of the cell-ordered position array that
// make up its 27-cell neighbourhood: one (dz, dy) stencil row per range pair, rows in the order dz-major, dy-minor,
// within a row the x run [c0-1, c0+1] (one range) or, where the x stencil wraps around the periodic box, the low run
// followed by the wrapped remainder (two ranges).  Both kernels derive the layout from (cell_start, GridParams) with
// the function below, so a TILE-LOCAL index written by the build (16 bits) addresses the same atom in the tile the
// force kernel stages.
#pragma once
#include "common.cuh"

struct TileRange { uint32_t src, cnt, off; };

struct TilePlan {
    TileRange r0, r1;   // this lane's two ranges (lanes 0..8 own a stencil row each; empty on lanes >= 9)
    uint32_t m;         // atoms in the tile (warp-uniform)
    uint32_t self_off;  // tile index of the cell's first atom a0 (warp-uniform)
    int wrap;           // 0: the stencil never wraps; 1: wraps, >= 3 cells per axis; 2: wraps, tiny grid (exact path)
};

#ifndef MC_HOST_SHIM
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// the same wait with a suspend-time hint: a waiting warp is parked by the hardware for up to `ns` instead of coming back to
// re-issue the test every few hundred cycles (measured in pair_tile.cu: 7.5 M retries = 13 % of all issued instructions)
__device__ __forceinline__ void mbar_wait_parked(uint64_t *bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
            : "memory");
    } while (!ok);
}

// generic-proxy writes (a peer GPU's stores made visible by an acquire, this warp's own shared-memory stores) ordered
// before the async-proxy reads of a following bulk copy
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
#else
// host build: the same operations on the stand-in's threads / fibers (tests/cpp/shim_*/)
inline void mbar_init(uint64_t *bar, uint32_t count) { shim_mbar_init(bar, count); }
inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { shim_mbar_arrive(bar, bytes); }
inline void mbar_arrive(uint64_t *bar) { shim_mbar_arrive(bar, 0); }
inline void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) { shim_bulk_copy(dst, src, bytes, bar); }
inline void mbar_wait(uint64_t *bar, uint32_t parity) { shim_mbar_wait(bar, parity); }
inline void mbar_wait_parked(uint64_t *bar, uint32_t parity, uint32_t) { shim_mbar_wait(bar, parity); }
inline void fence_proxy_async() {}
#endif

// Layout of the tile of cell c (a0 = its first slot).  Called by all 32 lanes of a warp; results in P (r0 / r1 per lane,
// the rest warp-uniform).
__device__ __forceinline__ void tile_plan(const GridParams &g, const uint32_t *__restrict__ cell_start, int c, uint32_t a0,
                                          int lane, TilePlan &P) {
    TileRange r0 = {0u, 0u, 0u}, r1 = {0u, 0u, 0u};
    const int c0 = c % g.nc[0], c1 = (c / g.nc[0]) % g.nc[1], c2 = c / (g.nc[0] * g.nc[1]);
    if (lane < 9) {  // one (dz, dy) stencil row per lane -> up to two contiguous slot ranges
        const int dy = lane % 3 - 1, dz = lane / 3 - 1;
        const int lo_y = (g.nc[1] >= 3 || !g.periodic) ? -1 : 0, hi_y = (g.nc[1] >= 2 || !g.periodic) ? 1 : 0;
        const int lo_z = (g.nc[2] >= 3 || !g.z_ring) ? -1 : 0, hi_z = (g.nc[2] >= 2 || !g.z_ring) ? 1 : 0;
        int ky = c1 + dy, kz = c2 + dz;
        bool ok = dy >= lo_y && dy <= hi_y && dz >= lo_z && dz <= hi_z;
        if (g.periodic) ky = (ky + g.nc[1]) % g.nc[1];
        else if (ky < 0 || ky >= g.nc[1]) ok = false;
        if (g.z_ring) kz = (kz + g.nc[2]) % g.nc[2];
        else if (kz < 0 || kz >= g.nc[2]) ok = false;
        if (ok) {
            const int rowbase = (kz * g.nc[1] + ky) * g.nc[0];
            int x0, x1, y0 = 0, y1 = -1;  // second run empty unless the x stencil wraps
            if (!g.periodic) { x0 = max(c0 - 1, 0); x1 = min(c0 + 1, g.nc[0] - 1); }
            else if (g.nc[0] < 3) { x0 = 0; x1 = g.nc[0] - 1; }
            else if (c0 == 0) { x0 = 0; x1 = 1; y0 = y1 = g.nc[0] - 1; }
            else if (c0 == g.nc[0] - 1) { x0 = 0; x1 = 0; y0 = g.nc[0] - 2; y1 = g.nc[0] - 1; }
            else { x0 = c0 - 1; x1 = c0 + 1; }
            r0.src = cell_start[rowbase + x0];
            r0.cnt = cell_start[rowbase + x1 + 1] - r0.src;
            if (y1 >= y0) {
                r1.src = cell_start[rowbase + y0];
                r1.cnt = cell_start[rowbase + y1 + 1] - r1.src;
            }
        }
    }
    // exclusive prefix of the range lengths across lanes (lane order = range order)
    const uint32_t mine = r0.cnt + r1.cnt;
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
        const uint32_t v = __shfl_up_sync(MC_FULL_MASK, inc, d);
        if (lane >= d) inc += v;
    }
    r0.off = inc - mine;
    r1.off = r0.off + r0.cnt;
    P.m = __shfl_sync(MC_FULL_MASK, inc, 8);
    // the own cell lives in stencil row (dz, dy) = (0, 0) = lane 4: in range r0 unless it is the
    // wrapped remainder r1 (c0 == nc0-1 with a wrapping x stencil)
    uint32_t so = 0;
    if (lane == 4) so = (a0 >= r0.src && a0 < r0.src + r0.cnt) ? r0.off + (a0 - r0.src) : r1.off + (a0 - r1.src);
    P.self_off = __shfl_sync(MC_FULL_MASK, so, 4);
    // the minimum image is decided on the GLOBAL cell coordinates (a decomposed rank sees the seam of the
    // periodic box only in the layers next to it)
    const int c2g = (c2 + g.kz_off) % g.ncz_global;
    const bool roomy = g.nc[0] >= 3 && g.nc[1] >= 3 && g.ncz_global >= 3;
    const bool interior = g.periodic && roomy && c0 >= 1 && c0 <= g.nc[0] - 2 && c1 >= 1 && c1 <= g.nc[1] - 2 &&
                          c2g >= 1 && c2g <= g.ncz_global - 2;
    P.wrap = (g.periodic && !interior) ? (roomy ? 1 : 2) : 0;
    P.r0 = r0;
    P.r1 = r1;
}
