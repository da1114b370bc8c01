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

// Two listed pairs at once in packed fp32 (add / mul / fma.rn.f32x2, sm_100): the differences are formed in scalar
// registers (each LDS.128 delivers one atom's x, y, z in consecutive registers, a packed operand needs the SAME component
// of the two atoms side by side), everything after them runs two pairs per instruction -- half the issue slots of the
// arithmetic that bounds pair_tile.cu.  LJ only (the Coulomb forms and the energy row sum use pair_term).
//   d = x_j - x_i (sign flipped: the caller negates the accumulated sums once per row; the squares, and with them the
//   exact-fp32 cutoff mask, are unchanged);  w3 = (1/r^2)^3;  |F|/r = (c12 w3 - c6) w3 / r^2  with c12 = 48 eps sigma^12,
//   c6 = 24 eps sigma^6 (the same quantity as 24 eps s6 (2 s6 - 1) / r^2 above, two multiplications shorter).
// ok1 == false: the second pair is padding (odd row length): its cutoff is set below zero.
struct Acc2 { float2 fx, fy, fz; };

#ifndef MC_HOST_SHIM
// inline PTX with explicit .rn: never contracted into an fma (the __fmul2_rn / __fadd2_rn intrinsics of CUDA 12.9 are --
// ptxas fused mul + add of the squared distance into FFMA2, which changes the cutoff mask in the last bit)
__device__ __forceinline__ unsigned long long mc_pack2(float2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 mc_unpack2(unsigned long long r) {
    float2 a;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
    return a;
}
__device__ __forceinline__ float2 mc_mul2(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(mc_pack2(a)), "l"(mc_pack2(b)));
    return mc_unpack2(r);
}
__device__ __forceinline__ float2 mc_add2(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(mc_pack2(a)), "l"(mc_pack2(b)));
    return mc_unpack2(r);
}
__device__ __forceinline__ float2 mc_fma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(mc_pack2(a)), "l"(mc_pack2(b)), "l"(mc_pack2(c)));
    return mc_unpack2(r);
}
#else
inline float2 mc_mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 mc_add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float2 mc_fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif

template <bool WRAP>
__device__ __forceinline__ void pair_term2_lj(const float4 xi, const float4 x0, const float4 x1, const float2 c12, const float2 c6n,
                                              const NbParams &p, const float rc2_0, const float rc2_1, Acc2 &a) {
    float2 dx = make_float2(x0.x - xi.x, x1.x - xi.x), dy = make_float2(x0.y - xi.y, x1.y - xi.y),
           dz = make_float2(x0.z - xi.z, x1.z - xi.z);
    if (WRAP) {
        dx.x = __fmaf_rn(-rintf(dx.x * p.inv_ext[0]), p.ext[0], dx.x); dx.y = __fmaf_rn(-rintf(dx.y * p.inv_ext[0]), p.ext[0], dx.y);
        dy.x = __fmaf_rn(-rintf(dy.x * p.inv_ext[1]), p.ext[1], dy.x); dy.y = __fmaf_rn(-rintf(dy.y * p.inv_ext[1]), p.ext[1], dy.y);
        dz.x = __fmaf_rn(-rintf(dz.x * p.inv_ext[2]), p.ext[2], dz.x); dz.y = __fmaf_rn(-rintf(dz.y * p.inv_ext[2]), p.ext[2], dz.y);
    }
    // the oracle's fp32 expression ((dx*dx)+(dy*dy))+(dz*dz), no contraction.  The squares are packed; the two additions are
    // scalar on purpose: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with explicit .rn (checked in
    // SASS), which would change the cutoff mask in the last bit, while it never fuses a scalar FADD into a packed product.
    const float2 sx = mc_mul2(dx, dx), sy = mc_mul2(dy, dy), sz = mc_mul2(dz, dz);
    const float2 r2 = make_float2(__fadd_rn(__fadd_rn(sx.x, sy.x), sz.x), __fadd_rn(__fadd_rn(sx.y, sy.y), sz.y));
    const float2 ir2 = make_float2(rcp_approx(r2.x), rcp_approx(r2.y));
    const float2 w = mc_mul2(ir2, ir2);
    const float2 w3 = mc_mul2(w, ir2);
    const float2 g = mc_fma2(c12, w3, c6n);   // c12 w3 - c6
    const float2 h = mc_mul2(w3, ir2);
    float2 f = mc_mul2(g, h);
    f.x = r2.x < rc2_0 ? f.x : 0.f;
    f.y = r2.y < rc2_1 ? f.y : 0.f;
    a.fx = mc_fma2(dx, f, a.fx);
    a.fy = mc_fma2(dy, f, a.fy);
    a.fz = mc_fma2(dz, f, a.fz);
}

}  // namespace
