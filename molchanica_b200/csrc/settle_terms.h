// settle_terms.h -- analytic SETTLE for rigid three-site water (Miyamoto & Kollman, J. Comput. Chem. 13, 952
// (1992)), SURVEY 8f row 2: the reference integrates its water rigidly (OPC / SETTLE, README.md:239,
// ui/panels/md.rs:362-371).  Written once for device and host like bonded_terms.h: settle.cu runs it on the
// GPU, tests/cpp/settle_math_host.cpp compiles the same function with g++ and tests/test_settle_cpu.py checks
// it against a converged fp64 SHAKE (the two solve the same equations) and the invariants (bond lengths,
// centre of mass).  Positions are handled relative to the oxygen's old position, in fp32.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MC_SETTLE_HD __host__ __device__ __forceinline__
#else
#define MC_SETTLE_HD inline
#endif

struct SettleParams {
    float m_o, m_h;   // masses
    float ra, rb, rc; // canonical triangle: O at (0, ra), H at (-+rc, -rb) about the centre of mass
    float d_hh;       // H-H distance (= 2 rc)
};

// d_oh, d_hh: constrained O-H and H-H distances.
MC_SETTLE_HD SettleParams mc_settle_params(float m_o, float m_h, float d_oh, float d_hh) {
    SettleParams p;
    p.m_o = m_o; p.m_h = m_h;
    p.rc = 0.5f * d_hh;
    const float height = sqrtf(d_oh * d_oh - p.rc * p.rc);  // O above the H-H line
    const float inv_m = 1.f / (m_o + 2.f * m_h);
    p.ra = height * 2.f * m_h * inv_m;  // O above the centre of mass
    p.rb = height * m_o * inv_m;        // H-H line below it
    p.d_hh = d_hh;
    return p;
}

// b0, c0: old H1, H2 positions relative to the old O position (a0 = 0).
// a1, b1, c1: unconstrained new positions of O, H1, H2 relative to the old O position.
// Writes the constrained new positions (same frame) into a3, b3, c3.
MC_SETTLE_HD void mc_settle(const SettleParams &p, const float b0[3], const float c0[3], const float a1[3], const float b1[3],
                            const float c1[3], float a3[3], float b3[3], float c3[3]) {
    const float inv_m = 1.f / (p.m_o + 2.f * p.m_h);
    float com[3], xa1[3], xb1[3], xc1[3];
    for (int x = 0; x < 3; ++x) {
        com[x] = (a1[x] * p.m_o + (b1[x] + c1[x]) * p.m_h) * inv_m;
        xa1[x] = a1[x] - com[x]; xb1[x] = b1[x] - com[x]; xc1[x] = c1[x] - com[x];
    }
    // orthonormal frame: Z normal to the OLD triangle, X = a1 x Z, Y = Z x X
    float Z[3] = {b0[1] * c0[2] - b0[2] * c0[1], b0[2] * c0[0] - b0[0] * c0[2], b0[0] * c0[1] - b0[1] * c0[0]};
    float X[3] = {xa1[1] * Z[2] - xa1[2] * Z[1], xa1[2] * Z[0] - xa1[0] * Z[2], xa1[0] * Z[1] - xa1[1] * Z[0]};
    float Y[3] = {Z[1] * X[2] - Z[2] * X[1], Z[2] * X[0] - Z[0] * X[2], Z[0] * X[1] - Z[1] * X[0]};
    const float nx = 1.f / sqrtf(X[0] * X[0] + X[1] * X[1] + X[2] * X[2]);
    const float ny = 1.f / sqrtf(Y[0] * Y[0] + Y[1] * Y[1] + Y[2] * Y[2]);
    const float nz = 1.f / sqrtf(Z[0] * Z[0] + Z[1] * Z[1] + Z[2] * Z[2]);
    for (int x = 0; x < 3; ++x) { X[x] *= nx; Y[x] *= ny; Z[x] *= nz; }
#define MC_DOT(u, v) ((u)[0] * (v)[0] + (u)[1] * (v)[1] + (u)[2] * (v)[2])
    const float xb0d = MC_DOT(X, b0), yb0d = MC_DOT(Y, b0);
    const float xc0d = MC_DOT(X, c0), yc0d = MC_DOT(Y, c0);
    const float za1d = MC_DOT(Z, xa1);
    const float xb1d = MC_DOT(X, xb1), yb1d = MC_DOT(Y, xb1), zb1d = MC_DOT(Z, xb1);
    const float xc1d = MC_DOT(X, xc1), yc1d = MC_DOT(Y, xc1), zc1d = MC_DOT(Z, xc1);

    const float sinphi = za1d / p.ra;
    const float cosphi = sqrtf(fmaxf(1.f - sinphi * sinphi, 0.f));
    const float sinpsi = (zb1d - zc1d) / (2.f * p.rc * cosphi);
    const float cospsi = sqrtf(fmaxf(1.f - sinpsi * sinpsi, 0.f));

    const float ya2d = p.ra * cosphi;
    const float xb2d = -p.rc * cospsi;
    const float yb2d = -p.rb * cosphi - p.rc * sinpsi * sinphi;
    const float yc2d = -p.rb * cosphi + p.rc * sinpsi * sinphi;

    const float alpha = xb2d * (xb0d - xc0d) + yb0d * yb2d + yc0d * yc2d;
    const float beta = xb2d * (yc0d - yb0d) + xb0d * yb2d + xc0d * yc2d;
    const float gamma = xb0d * yb1d - xb1d * yb0d + xc0d * yc1d - xc1d * yc0d;
    const float al2be2 = alpha * alpha + beta * beta;
    const float sintheta = (alpha * gamma - beta * sqrtf(fmaxf(al2be2 - gamma * gamma, 0.f))) / al2be2;
    const float costheta = sqrtf(fmaxf(1.f - sintheta * sintheta, 0.f));

    const float xa3d = -ya2d * sintheta, ya3d = ya2d * costheta, za3d = za1d;
    const float xb3d = xb2d * costheta - yb2d * sintheta, yb3d = xb2d * sintheta + yb2d * costheta, zb3d = zb1d;
    const float xc3d = -xb2d * costheta - yc2d * sintheta, yc3d = -xb2d * sintheta + yc2d * costheta, zc3d = zc1d;
#undef MC_DOT
    for (int x = 0; x < 3; ++x) {
        a3[x] = com[x] + X[x] * xa3d + Y[x] * ya3d + Z[x] * za3d;
        b3[x] = com[x] + X[x] * xb3d + Y[x] * yb3d + Z[x] * zb3d;
        c3[x] = com[x] + X[x] * xc3d + Y[x] * yc3d + Z[x] * zc3d;
    }
}
