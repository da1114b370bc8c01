// Micro-benchmark: issue rate of packed fma.rn.f32x2 (FFMA2) against scalar FFMA on sm_100a.
// Each thread runs ILP independent dependency chains; reports fp32 lane-FMAs per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_scalar(float *out, int iters, float a, float b) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_packed(float *out, int iters, float a, float b) {
    unsigned long long x[ILP], A, B;
    asm("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float v = threadIdx.x * 1e-3f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(A), "l"(B));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_mix(float *out, int iters, float a, float b) {  // FADD2 + FMUL2 + FFMA2 mix like the pair kernel
    unsigned long long x[ILP], A, B;
    asm("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float v = threadIdx.x * 1e-3f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(B));
            asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(A));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, threads = 512, blocks = sms * 4, iters = 20000;
    float *out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    auto report = [&](const char *name, float ms, double lane_fmas_per_thread_iter, int ilp) {
        double total = (double)blocks * threads * iters * ilp * lane_fmas_per_thread_iter;
        printf("%-28s %8.3f ms  %7.2f T lane-op/s  (%.1f lane-ops/clk/SM at %d MHz nominal)\n", name, ms, total / ms * 1e-9,
               total / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000);
    };
#define RUN(K, NAME, LPI, ILP)                                        \
    K<ILP><<<blocks, threads>>>(out, 100, 1.0001f, 0.5f);             \
    cudaEventRecord(a);                                               \
    K<ILP><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);           \
    cudaEventRecord(b); cudaEventSynchronize(b);                      \
    { float ms; cudaEventElapsedTime(&ms, a, b); report(NAME, ms, LPI, ILP); }
    RUN(k_scalar, "FFMA  ilp4", 1.0, 4)
    RUN(k_scalar, "FFMA  ilp8", 1.0, 8)
    RUN(k_packed, "FFMA2 ilp4", 2.0, 4)
    RUN(k_packed, "FFMA2 ilp8", 2.0, 8)
    RUN(k_mix, "FADD2+FMUL2 ilp4", 4.0, 4)
    RUN(k_mix, "FADD2+FMUL2 ilp8", 4.0, 8)
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
