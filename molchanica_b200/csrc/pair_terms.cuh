// pair_terms.cuh -- the arithmetic of one listed pair, shared by the two force kernels (pair_force.cu: rows of global
// indices gathered through L1; pair_tile.cu: rows of tile-local indices gathered from a TMA-staged shared-memory tile).
// Follows reference src/cuda/util.cu: LJ :93-139, Coulomb :54-63, minimum image :65-71 (see pair_force.cu's header).
#pragma once
#include "common.cuh"

namespace {

#ifndef MC_HOST_SHIM
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#else  // tests/cpp/pair_kernel_host.cpp runs this file's kernels on the CPU: no PTX there
inline float rcp_approx(float x) { return 1.0f / x; }
inline float rsqrt_approx(float x) { return 1.0f / sqrtf(x); }
#endif

struct Acc { float fx, fy, fz, e; };

// One listed pair.  WRAP: apply the minimum image (only rows of atoms in boundary cells need it).
template <int COUL, bool WRAP, bool ENERGY>
__device__ __forceinline__ void pair_term(const float4 xi, const float4 xj, const float2 lj, const NbParams &p,
                                          const float rc2_lj, Acc &a) {
    float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
    if (WRAP) {
        // rintf(d * inv_ext) == rintf(d / ext) except within rounding of |d| = ext/2, where both
        // images are beyond any legal cutoff (rc + skin <= ext/2 is enforced at build time)
        dx = __fmaf_rn(-rintf(dx * p.inv_ext[0]), p.ext[0], dx);
        dy = __fmaf_rn(-rintf(dy * p.inv_ext[1]), p.ext[1], dy);
        dz = __fmaf_rn(-rintf(dz * p.inv_ext[2]), p.ext[2], dz);
    }
    // the oracle's fp32 expression, no fma contraction: both sides mask exactly the same pairs
    const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    // branch-free LJ: the arithmetic runs for every listed pair and the cutoff selects the result (one FSEL
    // instead of a predicated block that re-materialises its constants); lj_on == false arrives as rc2 < 0
    const float ir2 = rcp_approx(r2);
    const float s2 = lj.x * ir2;
    const float s6 = s2 * s2 * s2;
    const bool in_lj = r2 < rc2_lj;
    float f = in_lj ? lj.y * s6 * __fmaf_rn(2.f, s6, -1.f) * ir2 : 0.f;
    float e = 0.f;
    if (ENERGY) e = in_lj ? lj.y * (1.f / 6.f) * s6 * (s6 - 1.f) : 0.f;
    if (COUL != MC_COULOMB_NONE) {
        if (r2 < p.rc2_q) {
            const float qq = xi.w * xj.w;
            const float ir = rsqrt_approx(r2);
            if (COUL == MC_COULOMB_PLAIN) {
                f = __fmaf_rn(qq * ir, rcp_approx(r2 + MC_SOFTENING_SQ), f);
                if (ENERGY) e = __fmaf_rn(qq, ir, e);
            } else {
                const float r = r2 * ir;
                const float ar = p.alpha * r;
                const float erfc_ar = erfcf(ar);
                const float ex = __expf(-ar * ar);
                const float ir2 = ir * ir;
                // |F|/r = qq (erfc(ar)/r^2 + 2a/sqrt(pi) exp(-a^2 r^2)/r) / r
                f = __fmaf_rn(qq * ir, __fmaf_rn(erfc_ar, ir2, 2.f * p.alpha * MC_INV_SQRT_PI * ex * ir), f);
                if (ENERGY) e = __fmaf_rn(qq * erfc_ar, ir, e);
            }
        }
    }
    a.fx = __fmaf_rn(dx, f, a.fx);
    a.fy = __fmaf_rn(dy, f, a.fy);
    a.fz = __fmaf_rn(dz, f, a.fz);
    if (ENERGY) a.e += e;
}

}  // namespace
