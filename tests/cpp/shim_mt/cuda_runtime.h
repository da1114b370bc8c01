/*
 * Test infrastructure: a second stand-in for <cuda_runtime.h>, for kernels whose threads COOPERATE (warp shuffles,
 * votes, __syncthreads).  One block runs at a time; its threads are real OS threads with thread-local threadIdx, a
 * std::barrier per block for __syncthreads and one per warp for the shuffle / vote exchanges, so the kernel sources of
 * the hot path (pair_force.cu) run unchanged on a machine without a GPU.  Every thread of a warp must reach every
 * shuffle / vote (true of the kernels run this way).  `__shared__` statics are shared by the whole process, hence the
 * one-block-at-a-time rule; dynamic shared memory is a process-wide array the harness defines under the kernel's name.
 * Slow (thousands of barrier waits per block): meant for systems of a few hundred atoms.
 */
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <barrier>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __align__(n) __attribute__((aligned(n)))
#ifdef MC_SHIM_SHARED_STATIC
#define __shared__ static   // statically sized shared arrays: one copy for the process = for the one block that runs
#else
#define __shared__          // kernels with `extern __shared__`: the harness defines the array under the kernel's name
#endif

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct shim_dim3 { unsigned x, y, z; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

static thread_local shim_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
static shim_dim3 blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

struct ShimBlock {
    std::unique_ptr<std::barrier<>> block_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
    std::vector<uint64_t> xchg;  // one 64-bit slot per thread
};
static ShimBlock *shim_block = nullptr;

static inline void __syncthreads() { shim_block->block_bar->arrive_and_wait(); }

template <typename T>
static inline T shim_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    const unsigned t = threadIdx.x, w = t >> 5;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    shim_block->xchg[t] = bits;
    shim_block->warp_bar[w]->arrive_and_wait();
    const unsigned src = (w << 5) | ((unsigned)src_lane & 31u);
    T out = v;
    if (src < blockDim.x) memcpy(&out, &shim_block->xchg[src], sizeof(T));
    shim_block->warp_bar[w]->arrive_and_wait();
    return out;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int d) { return shim_exchange(v, (int)(threadIdx.x & 31) ^ d); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int lane) { return shim_exchange(v, lane); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) {
    const int lane = (int)(threadIdx.x & 31);
    T got = shim_exchange(v, lane >= d ? lane - d : lane);
    return lane >= d ? got : v;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    const unsigned t = threadIdx.x, w = t >> 5;
    shim_block->xchg[t] = pred ? 1u : 0u;
    shim_block->warp_bar[w]->arrive_and_wait();
    unsigned m = 0;
    for (unsigned l = 0; l < 32 && (w << 5 | l) < blockDim.x; ++l) m |= (unsigned)(shim_block->xchg[w << 5 | l] & 1u) << l;
    shim_block->warp_bar[w]->arrive_and_wait();
    return m;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }

#include <atomic>
// Rendezvous of the lanes named in `mask` only (the callers sit inside a divergent branch, so the 32-wide warp barrier
// cannot be used): a counting barrier whose size is popc(mask), one per warp.
struct ShimMaskBar { std::atomic<unsigned> count{0}, gen{0}; };
static ShimMaskBar shim_mask_bar[32];
static inline void shim_mask_wait(unsigned w, unsigned n) {
    ShimMaskBar &b = shim_mask_bar[w];
    const unsigned g = b.gen.load();
    if (b.count.fetch_add(1) + 1 == n) {
        b.count.store(0);
        b.gen.fetch_add(1);
    } else {
        while (b.gen.load() == g) std::this_thread::yield();
    }
}
static inline unsigned __match_any_sync(unsigned mask, unsigned v) {
    const unsigned t = threadIdx.x, w = t >> 5, n = (unsigned)__builtin_popcount(mask);
    shim_block->xchg[t] = v;
    shim_mask_wait(w, n);
    unsigned m = 0;
    for (unsigned l = 0; l < 32; ++l)
        if ((mask >> l & 1u) && (unsigned)shim_block->xchg[w << 5 | l] == v) m |= 1u << l;
    shim_mask_wait(w, n);
    return m;
}

#include <atomic>
static inline float atomicAdd(float *a, float v) {
    std::atomic_ref<float> r(*a);
    float o = r.load();
    while (!r.compare_exchange_weak(o, o + v)) {}
    return o;
}
static inline double atomicAdd(double *a, double v) {
    std::atomic_ref<double> r(*a);
    double o = r.load();
    while (!r.compare_exchange_weak(o, o + v)) {}
    return o;
}
static inline int atomicOr(int *a, int v) { return std::atomic_ref<int>(*a).fetch_or(v); }
static inline unsigned atomicOr(unsigned *a, unsigned v) { return std::atomic_ref<unsigned>(*a).fetch_or(v); }
static inline unsigned atomicAdd(unsigned *a, unsigned v) { return std::atomic_ref<unsigned>(*a).fetch_add(v); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
alignas(128) static unsigned char shim_dyn_smem[256 * 1024];  // MC_DYN_SHARED(T, name) of common.cuh points here
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline void __syncwarp(unsigned = 0xffffffffu) { shim_block->warp_bar[threadIdx.x >> 5]->arrive_and_wait(); }
static inline int atomicMin(int *a, int v) { std::atomic_ref<int> r(*a); int o = r.load(); while (o > v && !r.compare_exchange_weak(o, v)) {} return o; }
static inline int atomicMax(int *a, int v) { std::atomic_ref<int> r(*a); int o = r.load(); while (o < v && !r.compare_exchange_weak(o, v)) {} return o; }
static inline unsigned atomicMin(unsigned *a, unsigned v) { std::atomic_ref<unsigned> r(*a); unsigned o = r.load(); while (o > v && !r.compare_exchange_weak(o, v)) {} return o; }
static inline unsigned atomicMax(unsigned *a, unsigned v) { std::atomic_ref<unsigned> r(*a); unsigned o = r.load(); while (o < v && !r.compare_exchange_weak(o, v)) {} return o; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
#define __expf(x) expf(x)
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }

typedef void *cudaStream_t;
typedef int cudaError_t;
static inline int cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
enum { cudaSuccess = 0 };
struct cudaFuncAttributes {};
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 1; return cudaSuccess; }

// Runs `body` (a call of the kernel with its arguments bound) for every thread of every block, one block at a time.
static inline void shim_launch(unsigned blocks, unsigned threads, const std::function<void()> &body) {
    gridDim = {blocks, 1, 1};
    blockDim = {threads, 1, 1};
    for (unsigned b = 0; b < blocks; ++b) {
        ShimBlock blk;
        blk.block_bar = std::make_unique<std::barrier<>>((ptrdiff_t)threads);
        const unsigned warps = (threads + 31) / 32;
        for (unsigned w = 0; w < warps; ++w) {
            const unsigned in_warp = (w + 1) * 32 <= threads ? 32 : threads - w * 32;
            blk.warp_bar.push_back(std::make_unique<std::barrier<>>((ptrdiff_t)in_warp));
        }
        blk.xchg.assign(threads, 0);
        shim_block = &blk;
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < threads; ++t)
            pool.emplace_back([&, t] {
                threadIdx = {t, 0, 0};
                blockIdx = {b, 0, 0};
                body();
            });
        for (auto &th : pool) th.join();
    }
    shim_block = nullptr;
}
