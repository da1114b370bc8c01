// bonded_terms.h -- per-term arithmetic of the bonded force field ("bonded forces on device"),
// written once for device and host: bonded.cu evaluates it on the GPU, tests/cpp/bonded_math_check.cpp compiles
// the very same functions with g++ and checks F = -dE/dx by central differences, so the arithmetic is verified
// on a machine without a GPU.  Amber functional forms (the reference's force field, README.md:234-241):
//   bond      E = k (r - r0)^2                        (no 1/2: Amber "RK")
//   angle     E = k (theta - theta0)^2                (Amber "TK"), theta in radians
//   dihedral  E = pk (1 + cos(n phi - gamma))         (Amber "PK" = Vn/2 / divider, "PN", "PHASE"),
//             phi by the IUPAC convention (sign of r_ij . (r_kj x r_kl))
// Forces follow the standard derivations (dihedral: Bekker et al. as used by GROMACS do_dih_fup).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MC_HD __host__ __device__ __forceinline__
#else
#define MC_HD inline
#endif

// d = r_i - r_j (minimum image applied by the caller).  Writes the force on i (the force on j is its negative).
MC_HD float mc_bond_term(const float d[3], float k, float r0, float f_i[3]) {
    const float r = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const float dr = r - r0;
    const float s = -2.f * k * dr / r;
    f_i[0] = d[0] * s; f_i[1] = d[1] * s; f_i[2] = d[2] * s;
    return k * dr * dr;
}

// a = r_i - r_j, b = r_k - r_j (j is the vertex).  Writes the forces on i and k; the force on j is -(f_i + f_k).
MC_HD float mc_angle_term(const float a[3], const float b[3], float k, float theta0, float f_i[3], float f_k[3]) {
    const float aa = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], bb = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
    const float ab = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    const float inv = 1.f / sqrtf(aa * bb);
    float c = ab * inv;
    c = fminf(1.f, fmaxf(-1.f, c));
    const float theta = acosf(c);
    const float dth = theta - theta0;
    const float s = sqrtf(fmaxf(1.f - c * c, 1e-12f));
    // dE/dtheta = 2 k dth;  dtheta/da = -(b/(|a||b|) - c a/|a|^2) / sin(theta)
    const float g = 2.f * k * dth / s;
    const float ca = c / aa, cb = c / bb;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int x = 0; x < 3; ++x) {
        f_i[x] = g * (b[x] * inv - ca * a[x]);
        f_k[x] = g * (a[x] * inv - cb * b[x]);
    }
    return k * dth * dth;
}

// r_ij = r_i - r_j, r_kj = r_k - r_j, r_kl = r_k - r_l.  Writes the forces on i, j, k, l.
MC_HD float mc_dihedral_term(const float r_ij[3], const float r_kj[3], const float r_kl[3], float pk, float n_per, float gamma,
                             float f_i[3], float f_j[3], float f_k[3], float f_l[3]) {
    const float m[3] = {r_ij[1] * r_kj[2] - r_ij[2] * r_kj[1], r_ij[2] * r_kj[0] - r_ij[0] * r_kj[2],
                        r_ij[0] * r_kj[1] - r_ij[1] * r_kj[0]};
    const float n[3] = {r_kj[1] * r_kl[2] - r_kj[2] * r_kl[1], r_kj[2] * r_kl[0] - r_kj[0] * r_kl[2],
                        r_kj[0] * r_kl[1] - r_kj[1] * r_kl[0]};
    const float mm = m[0] * m[0] + m[1] * m[1] + m[2] * m[2], nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    const float kj2 = r_kj[0] * r_kj[0] + r_kj[1] * r_kj[1] + r_kj[2] * r_kj[2];
    const float mn = m[0] * n[0] + m[1] * n[1] + m[2] * n[2];
    // phi = atan2(|r_kj| r_ij . n, m . n): the IUPAC angle without the acos singularities at 0 and pi
    const float kj = sqrtf(kj2);
    const float phi = atan2f(kj * (r_ij[0] * n[0] + r_ij[1] * n[1] + r_ij[2] * n[2]), mn);
    const float arg = n_per * phi - gamma;
    const float e = pk * (1.f + cosf(arg));
    const float ddphi = -pk * n_per * sinf(arg);  // dE/dphi
    const float safe_mm = fmaxf(mm, 1e-12f), safe_nn = fmaxf(nn, 1e-12f);
    const float ai = -ddphi * kj / safe_mm, bl = ddphi * kj / safe_nn;
    const float p = (r_ij[0] * r_kj[0] + r_ij[1] * r_kj[1] + r_ij[2] * r_kj[2]) / kj2;
    const float q = (r_kl[0] * r_kj[0] + r_kl[1] * r_kj[1] + r_kl[2] * r_kj[2]) / kj2;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int x = 0; x < 3; ++x) {
        const float fi = ai * m[x], fl = bl * n[x];
        const float sv = p * fi - q * fl;
        f_i[x] = fi;
        f_j[x] = -(fi - sv);
        f_k[x] = -(fl + sv);
        f_l[x] = fl;
    }
    return e;
}
