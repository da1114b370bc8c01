// sort_scan.cu -- hand-written device prefix scan and stable LSD radix sort used by the cell-list
// build ("cell-list build as on-device radix-sort + prefix-scan", BASELINE.json north_star;
// SURVEY 8a row a3).  HBM-bound integer work: coalesced 16-byte loads, warp-shuffle scans,
// warp-aggregated ranking (__match_any_sync) instead of per-key atomics.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t up8(uint32_t v) { return (v + 7u) & ~7u; }

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(MC_FULL_MASK, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns the exclusive prefix and the
// block total through *total.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = warp_inclusive_scan(v, lane);
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0u;
        uint32_t si = warp_inclusive_scan(s, lane);
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = si - s;  // exclusive warp offsets
        if (lane == SCAN_THREADS / 32 - 1) *total = si;
    }
    __syncthreads();
    return inc - v + warp_sums[wid];
}

template <bool ALIGN8>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t *__restrict__ in, size_t n,
                                                                    uint32_t *__restrict__ block_sums) {
    const size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        size_t i = base + (size_t)k * SCAN_THREADS + threadIdx.x;  // coalesced
        if (i < n) s += ALIGN8 ? up8(in[i]) : in[i];
    }
    __shared__ uint32_t tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// one block: in-place exclusive scan of the block sums; the grand total goes to *total_out
__global__ void __launch_bounds__(1024) scan_spine_kernel(uint32_t *__restrict__ sums, size_t nb,
                                                           uint32_t *__restrict__ total_out) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t base = 0; base < nb; base += 1024) {
        size_t i = base + threadIdx.x;
        uint32_t v = i < nb ? sums[i] : 0u;
        uint32_t inc = warp_inclusive_scan(v, lane);
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t s = warp_sums[lane];
            uint32_t si = warp_inclusive_scan(s, lane);
            warp_sums[lane] = si - s;
        }
        __syncthreads();
        uint32_t excl = inc - v + warp_sums[wid] + carry;
        if (i < nb) sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

template <bool ALIGN8>
__global__ void __launch_bounds__(SCAN_THREADS) scan_down_kernel(const uint32_t *__restrict__ in,
                                                                  uint32_t *__restrict__ out, size_t n,
                                                                  const uint32_t *__restrict__ block_sums) {
    // thread t owns SCAN_ITEMS consecutive elements -> 64-byte contiguous reads/writes per thread
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
    if (base + SCAN_ITEMS <= n) {
        const uint4 *p = reinterpret_cast<const uint4 *>(in + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
            uint4 q = p[k];
            v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = base + k < n ? in[base + k] : 0u;
    }
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (ALIGN8) v[k] = up8(v[k]);
        s += v[k];
    }
    __shared__ uint32_t tot;
    uint32_t run = block_exclusive_scan(s, &tot) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        uint32_t t = v[k];
        v[k] = run;
        run += t;
    }
    if (base + SCAN_ITEMS <= n) {
        uint4 *p = reinterpret_cast<uint4 *>(out + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) p[k] = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (base + k < n) out[base + k] = v[k];
    }
}

// ---- radix sort -----------------------------------------------------------------------------
constexpr int RS_BITS = 8;
constexpr int RS_BINS = 1 << RS_BITS;
constexpr int RS_ITEMS = 16;               // keys per lane
constexpr int RS_WARP_TILE = 32 * RS_ITEMS;  // keys per warp (contiguous)
constexpr int RS_WARPS = 8;                // warps per block

// per-warp digit histogram -> hist[digit * n_warps + warp]
__global__ void __launch_bounds__(RS_WARPS * 32) radix_hist_kernel(const uint32_t *__restrict__ keys, size_t n,
                                                                    int shift, uint32_t *__restrict__ hist,
                                                                    uint32_t n_warps) {
    __shared__ uint32_t h[RS_WARPS][RS_BINS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * RS_WARPS + w;
    for (int b = lane; b < RS_BINS; b += 32) h[w][b] = 0;
    __syncwarp();
    if (gw < n_warps) {
        const size_t base = (size_t)gw * RS_WARP_TILE;
#pragma unroll 4
        for (int k = 0; k < RS_ITEMS; ++k) {
            size_t i = base + (size_t)k * 32 + lane;
            bool ok = i < n;
            uint32_t act = __ballot_sync(MC_FULL_MASK, ok);
            if (ok) {
                uint32_t d = (keys[i] >> shift) & (RS_BINS - 1);
                uint32_t peers = __match_any_sync(act, d);
                if ((peers & ((1u << lane) - 1u)) == 0) h[w][d] += __popc(peers);  // group leader
            }
            __syncwarp();
        }
        for (int b = lane; b < RS_BINS; b += 32) hist[(size_t)b * n_warps + gw] = h[w][b];
    }
}

// stable scatter: ranks are (scanned global offset of (digit, warp)) + running count in the warp
__global__ void __launch_bounds__(RS_WARPS * 32) radix_scatter_kernel(const uint32_t *__restrict__ keys_in,
                                                                       const uint32_t *__restrict__ vals_in,
                                                                       uint32_t *__restrict__ keys_out,
                                                                       uint32_t *__restrict__ vals_out, size_t n,
                                                                       int shift, const uint32_t *__restrict__ offs,
                                                                       uint32_t n_warps) {
    __shared__ uint32_t o[RS_WARPS][RS_BINS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * RS_WARPS + w;
    if (gw >= n_warps) return;
    for (int b = lane; b < RS_BINS; b += 32) o[w][b] = offs[(size_t)b * n_warps + gw];
    __syncwarp();
    const size_t base = (size_t)gw * RS_WARP_TILE;
    for (int k = 0; k < RS_ITEMS; ++k) {
        size_t i = base + (size_t)k * 32 + lane;
        bool ok = i < n;
        uint32_t act = __ballot_sync(MC_FULL_MASK, ok);
        uint32_t key = 0, val = 0, d = 0, peers = 0, rank = 0;
        if (ok) {
            key = keys_in[i];
            val = vals_in[i];
            d = (key >> shift) & (RS_BINS - 1);
            peers = __match_any_sync(act, d);
            rank = __popc(peers & ((1u << lane) - 1u));
            uint32_t pos = o[w][d] + rank;
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
        __syncwarp();
        if (ok && rank == 0) o[w][d] += __popc(peers);
        __syncwarp();
    }
}

}  // namespace

size_t scan_scratch_elems(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 2; }

void exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n, int align8, uint32_t *scratch,
                        cudaStream_t st, int64_t *launches) {
    const size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb == 0) {
        cudaMemsetAsync(out, 0, sizeof(uint32_t), st);
        return;
    }
    if (align8) MC_LAUNCH(scan_reduce_kernel<true>, (unsigned)nb, SCAN_THREADS, 0, st, in, n, scratch);
    else MC_LAUNCH(scan_reduce_kernel<false>, (unsigned)nb, SCAN_THREADS, 0, st, in, n, scratch);
    MC_LAUNCH(scan_spine_kernel, 1, 1024, 0, st, scratch, nb, out + n);
    if (align8) MC_LAUNCH(scan_down_kernel<true>, (unsigned)nb, SCAN_THREADS, 0, st, in, out, n, scratch);
    else MC_LAUNCH(scan_down_kernel<false>, (unsigned)nb, SCAN_THREADS, 0, st, in, out, n, scratch);
    if (launches) *launches += 3;
}

static size_t radix_warps(size_t n) { return (n + RS_WARP_TILE - 1) / RS_WARP_TILE; }

size_t radix_scratch_elems(size_t n) {
    size_t table = radix_warps(n) * RS_BINS + 8;
    return table + scan_scratch_elems(table) + 8;
}

int radix_sort_pairs(uint32_t *keys[2], uint32_t *vals[2], size_t n, int bits, uint32_t *scratch,
                     cudaStream_t st, int64_t *launches) {
    if (n == 0) return 0;
    const uint32_t nw = (uint32_t)radix_warps(n);
    const size_t table = (size_t)nw * RS_BINS;
    uint32_t *hist = scratch;
    uint32_t *scan_scratch = scratch + ((table + 8) & ~(size_t)3);
    const unsigned blocks = div_up(nw, RS_WARPS);
    int cur = 0;
    for (int shift = 0; shift < bits; shift += RS_BITS) {
        MC_LAUNCH(radix_hist_kernel, blocks, RS_WARPS * 32, 0, st, keys[cur], n, shift, hist, nw);
        exclusive_scan_u32(hist, hist, table, 0, scan_scratch, st, launches);
        MC_LAUNCH(radix_scatter_kernel, blocks, RS_WARPS * 32, 0, st, keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, shift,
                  hist, nw);
        if (launches) *launches += 2;
        cur ^= 1;
    }
    return cur;
}
