/*
 * Test infrastructure: a stand-in for <cuda_runtime.h> that lets g++ compile the KERNEL SOURCES of
 * molchanica_b200/csrc/{bonded,settle,thermostat,pme}.cu unchanged and run them on the CPU, one thread at a time
 * (blockDim = 1, blockIdx.x = the global thread index, gridDim.x = the thread count).  Warp shuffles return 0, so a
 * "warp reduction" leaves each thread with its own value and every thread is lane 0 of its own warp: reductions
 * that end in an atomicAdd by lane 0 stay exact.  What this cannot show: data races between threads, and anything
 * about cuFFT.  Kernels that synchronise threads (__syncthreads) or carry inline PTX are not run this way.
 */
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __shared__ static          // blocks are one thread wide here: kernels that cooperate are compiled, not run
static inline void __syncthreads() {}

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct shim_dim3 { unsigned x, y, z; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

static shim_dim3 blockIdx = {0, 0, 0}, gridDim = {1, 1, 1};
static const shim_dim3 threadIdx = {0, 0, 0}, blockDim = {1, 1, 1};

static inline float atomicAdd(float *a, float v) { float o = *a; *a += v; return o; }
static inline double atomicAdd(double *a, double v) { double o = *a; *a += v; return o; }
static inline int atomicAdd(int *a, int v) { int o = *a; *a += v; return o; }
template <typename T> static inline T __shfl_xor_sync(unsigned, T, int) { return T(0); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }

typedef void *cudaStream_t;
typedef int cudaError_t;
