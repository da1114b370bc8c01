/*
 * cuda_host_shim.h -- the handful of CUDA names the reference's src/cuda/{cuda,util}.cu use,
 * defined for a plain g++ host compile so the UNMODIFIED reference sources can be built into
 * oracle/_ref/libref_cuda.so and executed on the CPU (one "thread": blockIdx = threadIdx = 0,
 * blockDim = gridDim = 1, so every grid-stride loop walks all elements serially).
 * Test infrastructure only; contains no reference code.
 */
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline

struct float3 { float x, y, z; };
struct uint3 { unsigned x, y, z; };
struct shim_dim3 { unsigned x, y, z; };

static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }

static const shim_dim3 blockIdx = {0, 0, 0};
static const shim_dim3 threadIdx = {0, 0, 0};
static const shim_dim3 blockDim = {1, 1, 1};
static const shim_dim3 gridDim = {1, 1, 1};

static inline float atomicAdd(float *addr, float v) { float old = *addr; *addr += v; return old; }
