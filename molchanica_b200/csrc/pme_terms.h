// pme_terms.h -- arithmetic of smooth particle-mesh Ewald (Essmann et al., J. Chem. Phys. 103, 8577 (1995)),
// the reference's long-range electrostatics is SPME (README.md:240, crate `ewald`, Cargo.toml:30).
// Shared by device (pme.cu) and host (tests/cpp/pme_math_host.cpp) like bonded_terms.h, so that the spline
// weights, the grid addressing and the influence function are verified on a machine without a GPU against the
// numpy restatement (oracle/pme_oracle.py), which in turn is checked against the exact Ewald sum.
// Order-4 cardinal B-splines; charges carry sqrt(332.0522), so energies are kcal/mol with Coulomb constant 1.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MC_PME_HD __host__ __device__ __forceinline__
#else
#define MC_PME_HD inline
#endif

#define MC_PME_ORDER 4

// w = u - floor(u) in [0, 1).  th[t], dth[t]: weight and d/du of the weight of grid point floor(u) - 3 + t.
MC_PME_HD void mc_bspline4(float w, float th[4], float dth[4]) {
    const float v = 1.f - w;
    th[0] = v * v * v * (1.f / 6.f);
    th[1] = (3.f * w * w * w - 6.f * w * w + 4.f) * (1.f / 6.f);
    th[2] = (-3.f * w * w * w + 3.f * w * w + 3.f * w + 1.f) * (1.f / 6.f);
    th[3] = w * w * w * (1.f / 6.f);
    dth[0] = -0.5f * v * v;
    dth[1] = 0.5f * (3.f * w * w - 4.f * w);
    dth[2] = 0.5f * (-3.f * w * w + 2.f * w + 1.f);
    dth[3] = 0.5f * w * w;
}

// Scaled fractional coordinate of x on an axis with K grid points: u in [0, K), k0 = floor(u), w = u - k0.
MC_PME_HD void mc_pme_coord(float x, float lo, float inv_ext, int K, int *k0, float *w) {
    float s = (x - lo) * inv_ext;
    s -= floorf(s);                 // [0, 1)
    float u = s * (float)K;
    int k = (int)u;
    if (k >= K) { k = K - 1; u = (float)K; }  // s rounded up to 1
    *k0 = k;
    *w = u - (float)k;
}

// Grid index of spline point t of an atom whose base index is k0: (k0 - 3 + t) mod K.
MC_PME_HD int mc_pme_wrap(int k0, int t, int K) {
    int k = k0 - 3 + t;
    return k < 0 ? k + K : k;
}

// |b(m)|^-2 denominator of Essmann eq. 4.4 for order 4: |1/6 + 4/6 e^{i t} + 1/6 e^{2 i t}|^2 = (2/3 + cos(t)/3)^2
MC_PME_HD double mc_pme_bmod4(int m, int K) {
    const double t = 6.283185307179586 * (double)m / (double)K;
    const double c = 2.0 / 3.0 + cos(t) / 3.0;
    return c * c;
}

// Influence function B(m) C(m) of reciprocal vector (m1/L1, m2/L2, m3/L3), m folded to (-K/2, K/2]:
// exp(-pi^2 m^2 / alpha^2) / (pi V m^2) / (bmod1 bmod2 bmod3); 0 for m = 0.
MC_PME_HD float mc_pme_influence(int i1, int i2, int i3, int K1, int K2, int K3, const float inv_ext[3], float inv_vol_pi,
                                 float pi2_over_alpha2, float bm1, float bm2, float bm3) {
    const int m1 = i1 > K1 / 2 ? i1 - K1 : i1, m2 = i2 > K2 / 2 ? i2 - K2 : i2, m3 = i3 > K3 / 2 ? i3 - K3 : i3;
    if (m1 == 0 && m2 == 0 && m3 == 0) return 0.f;
    const float h1 = (float)m1 * inv_ext[0], h2 = (float)m2 * inv_ext[1], h3 = (float)m3 * inv_ext[2];
    const float msq = h1 * h1 + h2 * h2 + h3 * h3;
    return expf(-pi2_over_alpha2 * msq) * inv_vol_pi / (msq * bm1 * bm2 * bm3);
}

// |m|^2 of the same reciprocal vector (for the virial: W_rec = sum_m E(m) (1 - 2 pi^2 m^2 / alpha^2), the trace of the
// tensor of Essmann et al. eq. 2.7 = -dE_rec/d(lambda) under a uniform scaling of box and positions)
MC_PME_HD float mc_pme_msq(int i1, int i2, int i3, int K1, int K2, int K3, const float inv_ext[3]) {
    const int m1 = i1 > K1 / 2 ? i1 - K1 : i1, m2 = i2 > K2 / 2 ? i2 - K2 : i2, m3 = i3 > K3 / 2 ? i3 - K3 : i3;
    const float h1 = (float)m1 * inv_ext[0], h2 = (float)m2 * inv_ext[1], h3 = (float)m3 * inv_ext[2];
    return h1 * h1 + h2 * h2 + h3 * h3;
}

// Correction for one excluded (or 1-4) pair: the reciprocal sum contains the full erf(alpha r)/r interaction of
// every pair; excluded pairs must not have it.  d = r_i - r_j.  Returns the energy -qq erf(ar)/r and writes the
// force on i.
MC_PME_HD float mc_pme_excl_term(const float d[3], float qq, float alpha, float f_i[3]) {
    const float r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    const float r = sqrtf(r2);
    const float ar = alpha * r;
    const float erf_ar = erff(ar);
    // E = -qq erf(ar)/r ;  F_i = -dE/dr r_hat = qq (2a/sqrt(pi) exp(-a^2 r^2)/r - erf(ar)/r^2) r_hat
    const float fr = qq * (1.1283791670955126f * alpha * expf(-ar * ar) - erf_ar / r) / r2;
    f_i[0] = d[0] * fr; f_i[1] = d[1] * fr; f_i[2] = d[2] * fr;
    return -qq * erf_ar / r;
}
