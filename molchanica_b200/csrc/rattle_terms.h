// rattle_terms.h -- velocity stage of RATTLE (Andersen, J. Comput. Phys. 52, 24 (1983)) for the small constraint clusters
// of this engine: a rigid three-site water (3 constraints) or a heavy atom with <= 3 constrained hydrogens.  After the
// closing half kick of a velocity-Verlet step the velocities carry the components of the unconstrained forces ALONG the
// constrained bonds; they are removed by solving the (linear, <= 3 x 3) system
//     sum_l lambda_l (s_il / m_i - s_jl / m_j)(r_l . r_k) = r_k . (v_i - v_j)        for every constraint k = (i, j)
// exactly, and applying v_a -= (1 / m_a) sum_l s_al lambda_l r_l  (s_al = +1 / -1 when a is the first / second atom of
// constraint l).  Afterwards d/dt |r_ij| = 0 for every constrained pair, so kinetic energy, temperature and the 2 KE term
// of the pressure read between steps are those of the constrained system (round-1 advisor finding: +1.6 % KE without).
// Written once for device (fp32) and host / oracle (fp64), like settle_terms.h.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MC_RATTLE_HD __host__ __device__ __forceinline__
#else
#define MC_RATTLE_HD inline
#endif

// nc constraints (ci[k], cj[k]) over the atoms 0..3 of one cluster; r[a]: positions (any common origin, minimum image
// already applied), inv_m[a], v[a] in/out.
template <typename T>
MC_RATTLE_HD void mc_rattle_velocity(int nc, const int *ci, const int *cj, const T (*r)[3], const T *inv_m, T (*v)[3]) {
    T rk[3][3], b[3], A[3][3];
    for (int k = 0; k < 3; ++k) {
        b[k] = 0;
        for (int l = 0; l < 3; ++l) A[k][l] = k == l ? (T)1 : (T)0;
        for (int x = 0; x < 3; ++x) rk[k][x] = 0;
    }
    for (int k = 0; k < nc; ++k) {
        for (int x = 0; x < 3; ++x) {
            rk[k][x] = r[ci[k]][x] - r[cj[k]][x];
            b[k] += rk[k][x] * (v[ci[k]][x] - v[cj[k]][x]);
        }
    }
    for (int k = 0; k < nc; ++k)
        for (int l = 0; l < nc; ++l) {
            const T si = (T)((ci[k] == ci[l]) - (ci[k] == cj[l])), sj = (T)((cj[k] == ci[l]) - (cj[k] == cj[l]));
            A[k][l] = (si * inv_m[ci[k]] - sj * inv_m[cj[k]]) * (rk[l][0] * rk[k][0] + rk[l][1] * rk[k][1] + rk[l][2] * rk[k][2]);
        }
    // 3 x 3 solve by Cramer's rule (unused rows / columns are identity)
    const T c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1], c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2], c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    const T det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
    const T inv = (T)1 / det;
    T lam[3];
    lam[0] = (b[0] * c00 + A[0][1] * (A[1][2] * b[2] - b[1] * A[2][2]) + A[0][2] * (b[1] * A[2][1] - A[1][1] * b[2])) * inv;
    lam[1] = (A[0][0] * (b[1] * A[2][2] - A[1][2] * b[2]) + b[0] * c01 + A[0][2] * (A[1][0] * b[2] - b[1] * A[2][0])) * inv;
    lam[2] = (A[0][0] * (A[1][1] * b[2] - b[1] * A[2][1]) + A[0][1] * (b[1] * A[2][0] - A[1][0] * b[2]) + b[0] * c02) * inv;
    for (int l = 0; l < nc; ++l)
        for (int x = 0; x < 3; ++x) {
            v[ci[l]][x] -= inv_m[ci[l]] * lam[l] * rk[l][x];
            v[cj[l]][x] += inv_m[cj[l]] * lam[l] * rk[l][x];
        }
}
