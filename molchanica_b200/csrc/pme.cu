// pme.cu -- SPME reciprocal space on the device: order-4 B-spline charge spreading, cuFFT
// R2C, influence-function multiply (+ energy), cuFFT C2R, force interpolation; plus the erf correction of the
// excluded pairs.  Together with coulomb_mode = MC_COULOMB_ERFC (the erfc real-space term of the pair kernel) and
// the self term this is the reference's electrostatics (README.md:240).  cuFFT is the plain library FFT and is
// resolved with dlopen like NCCL; everything around it is hand-written and shares its arithmetic with the host
// tests through pme_terms.h.
// Roofline: spreading and interpolation are atomics / gathers on a grid that sits in L2 (64^3 floats = 1 MB);
// 64 grid points per atom -> N x 64 x 4 B of L2 traffic each; the FFTs are HBM-bound for grids beyond L2.
// STATUS: as bonded.cu -- arithmetic verified on the host (tests/test_pme_cpu.py), kernels not yet run on
// hardware (round-1 GPU budget spent).
#include "common.cuh"
#ifdef MC_HAVE_LAUNCH
#include <dlfcn.h>
#include <stdlib.h>
#endif

#include <string>
#include <vector>

#include "pme.cuh"
#include "pme_terms.h"

namespace {

#ifdef MC_HAVE_LAUNCH  // the serial stand-in of tests/cpp/shim/ compiles the kernels only
struct CufftApi {
    int (*Plan3d)(int *, int, int, int, int) = nullptr;
    int (*ExecR2C)(int, float *, float2 *) = nullptr;
    int (*ExecC2R)(int, float2 *, float *) = nullptr;
    int (*SetStream)(int, cudaStream_t) = nullptr;
    int (*Destroy)(int) = nullptr;
    bool ok = false;
    std::string err;
};

CufftApi &cufft_api() {
    static CufftApi api;
    if (api.ok || !api.err.empty()) return api;
    void *h = nullptr;
    // MOLCHANICA_CUFFT_LIB names the library explicitly (a non-default install; the host build of tests/cpp/host_lib/
    // points it at a plain-DFT stand-in with the same five entry points)
    const char *forced = getenv("MOLCHANICA_CUFFT_LIB");
    for (const char *name : {forced ? forced : "libcufft.so.11", "libcufft.so.12", "libcufft.so"}) {
        h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (h || forced) break;
    }
    if (!h) { api.err = std::string("cannot load libcufft: ") + dlerror(); return api; }
#define MC_SYM(field, name)                                                \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));     \
    if (!api.field) { api.err = std::string("libcufft lacks ") + name; return api; }
    MC_SYM(Plan3d, "cufftPlan3d")
    MC_SYM(ExecR2C, "cufftExecR2C")
    MC_SYM(ExecC2R, "cufftExecC2R")
    MC_SYM(SetStream, "cufftSetStream")
    MC_SYM(Destroy, "cufftDestroy")
#undef MC_SYM
    api.ok = true;
    return api;
}

constexpr int CUFFT_R2C_ = 0x2a, CUFFT_C2R_ = 0x2c;

#endif

struct PmeGeom {
    int K[3];
    float lo[3], inv_ext[3];
    float scale[3];  // K / ext: d(u)/d(x)
};

__global__ void __launch_bounds__(128) pme_spread_kernel(int n, const float4 *__restrict__ xyzq, const PmeGeom g,
                                                          float *__restrict__ grid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = xyzq[i];
    if (p.w == 0.f) return;
    int k0[3];
    float th[3][4], dth[4], w;
    mc_pme_coord(p.x, g.lo[0], g.inv_ext[0], g.K[0], &k0[0], &w); mc_bspline4(w, th[0], dth);
    mc_pme_coord(p.y, g.lo[1], g.inv_ext[1], g.K[1], &k0[1], &w); mc_bspline4(w, th[1], dth);
    mc_pme_coord(p.z, g.lo[2], g.inv_ext[2], g.K[2], &k0[2], &w); mc_bspline4(w, th[2], dth);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int ia = mc_pme_wrap(k0[0], a, g.K[0]);
        const float qa = p.w * th[0][a];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int ib = mc_pme_wrap(k0[1], b, g.K[1]);
            const float qab = qa * th[1][b];
            float *row = grid + ((size_t)ia * g.K[1] + ib) * g.K[2];
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(row + mc_pme_wrap(k0[2], c, g.K[2]), qab * th[2][c]);
        }
    }
}

// cgrid *= B C; energy = 1/2 sum B C |Q^|^2 with the Hermitian half counted twice where it stands for two modes
__global__ void __launch_bounds__(256) pme_convolve_kernel(int K1, int K2, int K3, float2 *__restrict__ cgrid,
                                                            const float *__restrict__ bm1, const float *__restrict__ bm2,
                                                            const float *__restrict__ bm3, float inv_e0, float inv_e1, float inv_e2,
                                                            float inv_vol_pi, float pi2_over_alpha2, double *__restrict__ energy,
                                                            int want_energy) {
    const int K3h = K3 / 2 + 1;
    const size_t total = (size_t)K1 * K2 * K3h;
    float e = 0.f, w = 0.f;
    const float inv_ext[3] = {inv_e0, inv_e1, inv_e2};
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i3 = (int)(idx % K3h);
        const int i2 = (int)((idx / K3h) % K2);
        const int i1 = (int)(idx / ((size_t)K3h * K2));
        const float bc = mc_pme_influence(i1, i2, i3, K1, K2, K3, inv_ext, inv_vol_pi, pi2_over_alpha2, bm1[i1], bm2[i2], bm3[i3]);
        float2 v = cgrid[idx];
        const float mult = (i3 == 0 || (K3 % 2 == 0 && i3 == K3 / 2)) ? 1.f : 2.f;
        const float em = 0.5f * mult * bc * (v.x * v.x + v.y * v.y);
        e += em;
        if (want_energy) w += em * (1.f - 2.f * pi2_over_alpha2 * mc_pme_msq(i1, i2, i3, K1, K2, K3, inv_ext));
        v.x *= bc; v.y *= bc;
        cgrid[idx] = v;
    }
    if (want_energy) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) e += __shfl_xor_sync(MC_FULL_MASK, e, d);
        if ((threadIdx.x & 31) == 0 && e != 0.f) atomicAdd(energy, (double)e);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(MC_FULL_MASK, w, d);
        if ((threadIdx.x & 31) == 0 && w != 0.f) atomicAdd(energy + 2, (double)w);  // energy[2]: reciprocal-space virial
    }
}

__global__ void __launch_bounds__(128) pme_gather_kernel(int n, const float4 *__restrict__ xyzq, const PmeGeom g,
                                                          const float *__restrict__ grid, float4 *__restrict__ force) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = xyzq[i];
    if (p.w == 0.f) return;
    int k0[3];
    float th[3][4], dth[3][4], w;
    mc_pme_coord(p.x, g.lo[0], g.inv_ext[0], g.K[0], &k0[0], &w); mc_bspline4(w, th[0], dth[0]);
    mc_pme_coord(p.y, g.lo[1], g.inv_ext[1], g.K[1], &k0[1], &w); mc_bspline4(w, th[1], dth[1]);
    mc_pme_coord(p.z, g.lo[2], g.inv_ext[2], g.K[2], &k0[2], &w); mc_bspline4(w, th[2], dth[2]);
    float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int ia = mc_pme_wrap(k0[0], a, g.K[0]);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int ib = mc_pme_wrap(k0[1], b, g.K[1]);
            const float *row = grid + ((size_t)ia * g.K[1] + ib) * g.K[2];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float phi = __ldg(row + mc_pme_wrap(k0[2], c, g.K[2]));
                fx += phi * dth[0][a] * th[1][b] * th[2][c];
                fy += phi * th[0][a] * dth[1][b] * th[2][c];
                fz += phi * th[0][a] * th[1][b] * dth[2][c];
            }
        }
    }
    // F = -q dE/dQ . dQ/dr ; only this thread touches force[i] in this kernel
    float4 f = force[i];
    f.x -= p.w * fx * g.scale[0];
    f.y -= p.w * fy * g.scale[1];
    f.z -= p.w * fz * g.scale[2];
    force[i] = f;
}

// one thread per atom: its excluded partners (both directions are listed, so no atomics; energy counted half)
__global__ void __launch_bounds__(128) pme_excl_kernel(int n, const float4 *__restrict__ xyzq, const int *__restrict__ orig,
                                                        const int *__restrict__ slot_of_orig, const int32_t *__restrict__ excl_start,
                                                        const int32_t *__restrict__ excl_idx, const NbParams p,
                                                        float4 *__restrict__ force, double *__restrict__ energy, int want_energy) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    float e = 0.f, w = 0.f;
    if (k < n) {
        const int oi = orig[k];
        const int e0 = excl_start[oi], e1 = excl_start[oi + 1];
        if (e1 > e0) {
            const float4 xi = xyzq[k];
            float fx = 0.f, fy = 0.f, fz = 0.f;
            for (int t = e0; t < e1; ++t) {
                const int j = slot_of_orig[excl_idx[t]];
                if (j < 0 || j == k) continue;
                const float4 xj = xyzq[j];
                float d[3] = {xi.x - xj.x, xi.y - xj.y, xi.z - xj.z}, f[3];
                if (p.periodic) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) d[a] -= rintf(d[a] * p.inv_ext[a]) * p.ext[a];
                }
                e += 0.5f * mc_pme_excl_term(d, xi.w * xj.w, p.alpha, f);
                w += 0.5f * (d[0] * f[0] + d[1] * f[1] + d[2] * f[2]);
                fx += f[0]; fy += f[1]; fz += f[2];
            }
            float4 f4 = force[k];
            f4.x += fx; f4.y += fy; f4.z += fz;
            force[k] = f4;
        }
    }
    if (want_energy) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) e += __shfl_xor_sync(MC_FULL_MASK, e, d);
        if ((threadIdx.x & 31) == 0 && e != 0.f) atomicAdd(energy, (double)e);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(MC_FULL_MASK, w, d);
        if ((threadIdx.x & 31) == 0 && w != 0.f) atomicAdd(energy + 2, (double)w);  // (s->energy + 1) + 2: virial of the correction
    }
}

}  // namespace

#ifdef MC_HAVE_LAUNCH  // the serial stand-in of tests/cpp/shim/ has no launcher
void pme_release(PmeState *s) {
    if (s->planned && cufft_api().ok) { cufft_api().Destroy(s->plan_r2c); cufft_api().Destroy(s->plan_c2r); }
    s->planned = false;
    if (s->grid) cudaFree(s->grid);
    if (s->cgrid) cudaFree(s->cgrid);
    for (int a = 0; a < 3; ++a) { if (s->bmod[a]) cudaFree(s->bmod[a]); s->bmod[a] = nullptr; }
    if (s->energy) cudaFree(s->energy);
    s->grid = nullptr; s->cgrid = nullptr; s->energy = nullptr;
    s->K[0] = s->K[1] = s->K[2] = 0;
}

int pme_configure(PmeState *s, int k1, int k2, int k3, cudaStream_t st, const char **msg) {
    pme_release(s);
    if (k1 == 0 && k2 == 0 && k3 == 0) return MC_OK;
    if (k1 < 8 || k2 < 8 || k3 < 8 || k1 > 2048 || k2 > 2048 || k3 > 2048) { *msg = "PME grid dimensions must lie in [8, 2048]"; return MC_E_INVALID; }
    CufftApi &api = cufft_api();
    if (!api.ok) { *msg = api.err.c_str(); return MC_E_CUDA; }
    const size_t nreal = (size_t)k1 * k2 * k3, ncplx = (size_t)k1 * k2 * (k3 / 2 + 1);
    if (cudaMalloc(&s->grid, nreal * sizeof(float)) != cudaSuccess || cudaMalloc(&s->cgrid, ncplx * sizeof(float2)) != cudaSuccess ||
        cudaMalloc(&s->energy, 4 * sizeof(double)) != cudaSuccess) {
        pme_release(s);
        *msg = "PME grid allocation failed";
        return MC_E_CUDA;
    }
    const int K[3] = {k1, k2, k3};
    for (int a = 0; a < 3; ++a) {
        std::vector<float> h((size_t)K[a]);
        for (int m = 0; m < K[a]; ++m) h[(size_t)m] = (float)mc_pme_bmod4(m, K[a]);
        if (cudaMalloc(&s->bmod[a], sizeof(float) * (size_t)K[a]) != cudaSuccess ||
            cudaMemcpy(s->bmod[a], h.data(), sizeof(float) * (size_t)K[a], cudaMemcpyHostToDevice) != cudaSuccess) {
            pme_release(s);
            *msg = "PME modulus upload failed";
            return MC_E_CUDA;
        }
    }
    if (api.Plan3d(&s->plan_r2c, k1, k2, k3, CUFFT_R2C_) != 0 || api.Plan3d(&s->plan_c2r, k1, k2, k3, CUFFT_C2R_) != 0 ||
        api.SetStream(s->plan_r2c, st) != 0 || api.SetStream(s->plan_c2r, st) != 0) {
        pme_release(s);
        *msg = "cufftPlan3d failed";
        return MC_E_CUDA;
    }
    s->planned = true;
    s->K[0] = k1; s->K[1] = k2; s->K[2] = k3;
    return MC_OK;
}

int pme_launch(PmeState *s, int n, const float4 *xyzq, const float lo[3], const float ext[3], float alpha, float4 *force,
               bool want_energy, cudaStream_t st, int64_t *launches, const char **msg) {
    if (!s->planned || n <= 0) return MC_OK;
    PmeGeom g;
    for (int a = 0; a < 3; ++a) {
        g.K[a] = s->K[a];
        g.lo[a] = lo[a];
        g.inv_ext[a] = 1.0f / ext[a];
        g.scale[a] = (float)s->K[a] / ext[a];
    }
    const size_t nreal = (size_t)s->K[0] * s->K[1] * s->K[2];
    cudaMemsetAsync(s->grid, 0, nreal * sizeof(float), st);
    if (want_energy) cudaMemsetAsync(s->energy, 0, 4 * sizeof(double), st);
    MC_LAUNCH(pme_spread_kernel, div_up((size_t)n, 128), 128, 0, st, n, xyzq, g, s->grid);
    CufftApi &api = cufft_api();
    if (api.ExecR2C(s->plan_r2c, s->grid, s->cgrid) != 0) { *msg = "cufftExecR2C failed"; return MC_E_CUDA; }
    const double vol = (double)ext[0] * ext[1] * ext[2];
    const float pi = 3.14159265358979f;
    MC_LAUNCH(pme_convolve_kernel, 592, 256, 0, st, s->K[0], s->K[1], s->K[2], s->cgrid, s->bmod[0], s->bmod[1], s->bmod[2], g.inv_ext[0],
                                            g.inv_ext[1], g.inv_ext[2], (float)(1.0 / (3.14159265358979323846 * vol)),
                                            pi * pi / (alpha * alpha), s->energy, want_energy ? 1 : 0);
    if (api.ExecC2R(s->plan_c2r, s->cgrid, s->grid) != 0) { *msg = "cufftExecC2R failed"; return MC_E_CUDA; }
    MC_LAUNCH(pme_gather_kernel, div_up((size_t)n, 128), 128, 0, st, n, xyzq, g, s->grid, force);
    *launches += 3;
    return MC_OK;
}

void pme_launch_exclusions(PmeState *s, int n, const float4 *xyzq, const int *orig, const int *slot_of_orig,
                           const int32_t *excl_start, const int32_t *excl_idx, const NbParams &p, float4 *force,
                           bool want_energy, cudaStream_t st, int64_t *launches) {
    if (!s->planned || n <= 0 || !excl_start) return;
    MC_LAUNCH(pme_excl_kernel, div_up((size_t)n, 128), 128, 0, st, n, xyzq, orig, slot_of_orig, excl_start, excl_idx, p, force,
                                                          s->energy + 1, want_energy ? 1 : 0);
    *launches += 1;
}
#endif
