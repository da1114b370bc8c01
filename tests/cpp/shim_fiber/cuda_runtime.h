/*
 * Test infrastructure: a third stand-in for <cuda_runtime.h>, complete enough to build the WHOLE library
 * (molchanica_b200/csrc/*.cu except comm.cu) for the host -- kernels, launchers and the C ABI of engine.cu -- so that
 * the parity tests written for the GPU can exercise every code path on a machine without one.
 *
 *   threads  : the threads of a block are FIBERS (user-space contexts, tests/cpp/shim_fiber/runtime.cpp) scheduled
 *              cooperatively on one OS thread: a warp shuffle costs 64 context switches of ~10 ns instead of 64 futex
 *              waits.  Blocks of one launch are spread over a small pool of OS threads.  Exited threads count as
 *              arrived at barriers, as on Volta and later.
 *   memory   : cudaMalloc = aligned host memory filled with 0xFF (NaN / -1: reads of never-written memory show up),
 *              copies are memcpy, streams and events are tokens, a launch has completed when the call returns.
 *   __shared__: `static thread_local` -- one copy per OS thread = per block in flight.
 *
 * What it cannot show: anything about timing, memory-model races between warps, cuFFT (tests/cpp/host_lib/ supplies a
 * plain DFT), NCCL / peer memory (comm.cu is replaced by a stub).  Not part of the product library.
 */
#pragma once
#include <math.h>
#include <sched.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <functional>

#define MC_HOST_LAUNCH 1  // common.cuh: launchers are compiled too (MC_LAUNCH -> shim_launch)

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local
#define __constant__ static

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct double3 { double x, y, z; };
struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct shim_dim3 { unsigned x, y, z; };
typedef shim_dim3 dim3;
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double3 make_double3(double x, double y, double z) { return double3{x, y, z}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

// ---- execution state (runtime.cpp) -------------------------------------------------------------------------
struct ShimThreadState {
    shim_dim3 tidx, bidx, bdim, gdim;
    unsigned char *dyn_smem;  // 256 KB per OS thread
};
extern thread_local ShimThreadState shim_ts;
#define threadIdx (shim_ts.tidx)
#define blockIdx (shim_ts.bidx)
#define blockDim (shim_ts.bdim)
#define gridDim (shim_ts.gdim)
#define shim_dyn_smem (shim_ts.dyn_smem)

void shim_launch(unsigned blocks, unsigned threads, const std::function<void()> &body);
static inline void shim_launch(shim_dim3 blocks, unsigned threads, const std::function<void()> &body) { shim_launch(blocks.x, threads, body); }
void shim_sync_block();                          // __syncthreads
void shim_sync_warp();                           // full-mask warp barrier (live lanes)
void shim_sync_mask(unsigned mask);              // barrier of the lanes named in mask
void shim_named_barrier(int id, unsigned n_threads);
void shim_yield_block();                         // inside a spin loop on something another warp / the copy engine provides
uint64_t *shim_xchg();                           // one 64-bit slot per thread of the running block

static inline void __syncthreads() { shim_sync_block(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { if (mask == 0xffffffffu) shim_sync_warp(); else shim_sync_mask(mask); }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline void __threadfence_block() {}

template <typename T>
static inline T shim_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    const unsigned t = threadIdx.x, w = t >> 5;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    uint64_t *x = shim_xchg();
    x[t] = bits;
    shim_sync_warp();
    const unsigned src = (w << 5) | ((unsigned)src_lane & 31u);
    T out = v;
    if (src < blockDim.x) memcpy(&out, &x[src], sizeof(T));
    shim_sync_warp();
    return out;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int d) { return shim_exchange(v, (int)(threadIdx.x & 31) ^ d); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int lane) { return shim_exchange(v, lane); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) {
    const int lane = (int)(threadIdx.x & 31);
    T got = shim_exchange(v, lane >= d ? lane - d : lane);
    return lane >= d ? got : v;
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) {
    const int lane = (int)(threadIdx.x & 31);
    T got = shim_exchange(v, lane + d < 32 ? lane + d : lane);
    return lane + d < 32 ? got : v;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    const unsigned t = threadIdx.x, w = t >> 5;
    uint64_t *x = shim_xchg();
    x[t] = pred ? 1u : 0u;
    __syncwarp(mask);
    unsigned m = 0;
    for (unsigned l = 0; l < 32 && (w << 5 | l) < blockDim.x; ++l)
        if (mask >> l & 1u) m |= (unsigned)(x[w << 5 | l] & 1u) << l;
    __syncwarp(mask);
    return m;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
static inline unsigned __match_any_sync(unsigned mask, unsigned v) {
    const unsigned t = threadIdx.x, w = t >> 5;
    uint64_t *x = shim_xchg();
    x[t] = v;
    shim_sync_mask(mask);
    unsigned m = 0;
    for (unsigned l = 0; l < 32; ++l)
        if ((mask >> l & 1u) && (unsigned)x[w << 5 | l] == v) m |= 1u << l;
    shim_sync_mask(mask);
    return m;
}

// ---- atomics (blocks of one launch run on several OS threads) ------------------------------------------------
template <typename T> static inline T shim_atomic_add_cas(T *a, T v) {
    std::atomic_ref<T> r(*a);
    T o = r.load();
    while (!r.compare_exchange_weak(o, o + v)) {}
    return o;
}
static inline float atomicAdd(float *a, float v) { return shim_atomic_add_cas(a, v); }
static inline double atomicAdd(double *a, double v) { return shim_atomic_add_cas(a, v); }
static inline int atomicAdd(int *a, int v) { return std::atomic_ref<int>(*a).fetch_add(v); }
static inline unsigned atomicAdd(unsigned *a, unsigned v) { return std::atomic_ref<unsigned>(*a).fetch_add(v); }
static inline unsigned long long atomicAdd(unsigned long long *a, unsigned long long v) { return std::atomic_ref<unsigned long long>(*a).fetch_add(v); }
static inline int atomicOr(int *a, int v) { return std::atomic_ref<int>(*a).fetch_or(v); }
static inline unsigned atomicOr(unsigned *a, unsigned v) { return std::atomic_ref<unsigned>(*a).fetch_or(v); }
static inline int atomicExch(int *a, int v) { return std::atomic_ref<int>(*a).exchange(v); }
static inline unsigned atomicExch(unsigned *a, unsigned v) { return std::atomic_ref<unsigned>(*a).exchange(v); }
static inline unsigned atomicCAS(unsigned *a, unsigned cmp, unsigned v) { std::atomic_ref<unsigned>(*a).compare_exchange_strong(cmp, v); return cmp; }
static inline int atomicCAS(int *a, int cmp, int v) { std::atomic_ref<int>(*a).compare_exchange_strong(cmp, v); return cmp; }
template <typename T> static inline T shim_atomic_min(T *a, T v) { std::atomic_ref<T> r(*a); T o = r.load(); while (o > v && !r.compare_exchange_weak(o, v)) {} return o; }
template <typename T> static inline T shim_atomic_max(T *a, T v) { std::atomic_ref<T> r(*a); T o = r.load(); while (o < v && !r.compare_exchange_weak(o, v)) {} return o; }
static inline int atomicMin(int *a, int v) { return shim_atomic_min(a, v); }
static inline int atomicMax(int *a, int v) { return shim_atomic_max(a, v); }
static inline unsigned atomicMin(unsigned *a, unsigned v) { return shim_atomic_min(a, v); }
static inline unsigned atomicMax(unsigned *a, unsigned v) { return shim_atomic_max(a, v); }

// ---- intrinsics ---------------------------------------------------------------------------------------------------
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline float __double2float_rn(double a) { return (float)a; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline long long __double_as_longlong(double d) { long long i; memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; memcpy(&d, &i, 8); return d; }
static inline float __saturatef(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
#define __expf(x) expf(x)
#define __logf(x) logf(x)
#define __cosf(x) cosf(x)
#define __sinf(x) sinf(x)
#define __sincosf(x, s, c) sincosf((x), (s), (c))
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t max(size_t a, size_t b) { return a > b ? a : b; }
static inline void __nanosleep(unsigned) { shim_yield_block(); sched_yield(); }  // a spin on memory another warp, block or PROCESS writes

// ---- host runtime (runtime.cpp) -------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 11, cudaErrorNotReady = 600 };
struct ShimStream;
struct ShimEvent;
typedef ShimStream *cudaStream_t;
typedef ShimEvent *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount, l2CacheSize;
    size_t totalGlobalMem, sharedMemPerBlockOptin;
};
cudaError_t shim_malloc(void **p, size_t bytes);
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return shim_malloc(reinterpret_cast<void **>(p), bytes); }
template <typename T> static inline cudaError_t cudaMallocHost(T **p, size_t bytes) { return shim_malloc(reinterpret_cast<void **>(p), bytes); }
template <typename T> static inline cudaError_t cudaHostAlloc(T **p, size_t bytes, unsigned) { return shim_malloc(reinterpret_cast<void **>(p), bytes); }
cudaError_t cudaFree(void *p);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st = nullptr);
cudaError_t cudaMemset(void *p, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t st = nullptr);
cudaError_t cudaSetDevice(int dev);
cudaError_t cudaGetDevice(int *dev);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int dev);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaGetLastError();
cudaError_t cudaPeekAtLastError();
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaStreamCreate(cudaStream_t *st);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *st, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t st);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaStreamQuery(cudaStream_t st);
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t ev, unsigned flags = 0);
cudaError_t cudaEventCreate(cudaEvent_t *ev);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *ev, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t ev);
cudaError_t cudaEventRecord(cudaEvent_t ev, cudaStream_t st = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t ev);
cudaError_t cudaEventQuery(cudaEvent_t ev);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 1; return cudaSuccess; }

// ---- peer memory between the processes of a decomposed host run (runtime.cpp): with MC_SHIM_SHARED_HEAP=1 "device"
// memory comes from a shared-memory arena per process, an IPC handle names (process, offset), opening it maps the
// exporter's arena.  Without the variable the calls fail and comm.cu falls back to the NCCL halo.
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1, cudaEnableDefault = 0, cudaErrorNotSupported = 801 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *base);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
cudaError_t cudaGetDriverEntryPoint(const char *name, void **fn, unsigned long long flags, cudaDriverEntryPointQueryResult *q = nullptr);

// ---- mbarrier / bulk copy / named barrier (tile_build.cu), cooperative versions ---------------------------------------------------
void shim_mbar_init(uint64_t *bar, uint32_t count);
void shim_mbar_arrive(uint64_t *bar, uint32_t tx_bytes);
void shim_bulk_copy(void *dst, const void *src, uint32_t bytes, uint64_t *bar);
void shim_mbar_wait(uint64_t *bar, uint32_t parity);
