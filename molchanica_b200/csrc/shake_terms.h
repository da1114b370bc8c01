// shake_terms.h -- SHAKE (Ryckaert, Ciccotti & Berendsen, J. Comput. Phys. 23, 327 (1977)) for bonds to hydrogen:
// one heavy atom with up to three hydrogens (CH, CH2, CH3, NH3+ ...), the constraint set the reference applies at
// 2 fs ("H-bond constraints", ui/panels/md.rs:362-371).  The constraints of one cluster couple only
// through the heavy atom, so one thread iterates them (Gauss-Seidel) to convergence.  Displacements go along the
// OLD bond vectors.  Shared by device (settle.cu) and host tests like settle_terms.h; positions are handled relative
// to the heavy atom's old position, in fp32.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MC_SHAKE_HD __host__ __device__ __forceinline__
#else
#define MC_SHAKE_HD inline
#endif

#define MC_SHAKE_MAX_H 3

// r0[k]: old vector heavy -> hydrogen k.  p0: new heavy position, p[k]: new hydrogen positions (relative to the OLD heavy
// position), updated in place.  inv_m0 / inv_m[k]: inverse masses, d[k]: constrained lengths, nh <= 3 hydrogens.
// Returns the number of sweeps used (max_iter + 1 = not converged).
MC_SHAKE_HD int mc_shake_cluster(int nh, const float r0[][3], float p0[3], float p[][3], float inv_m0, const float inv_m[], const float d[],
                                 float tol, int max_iter) {
    for (int it = 1; it <= max_iter; ++it) {
        float worst = 0.f;
        for (int k = 0; k < nh; ++k) {
            const float s[3] = {p[k][0] - p0[0], p[k][1] - p0[1], p[k][2] - p0[2]};
            const float ss = s[0] * s[0] + s[1] * s[1] + s[2] * s[2];
            const float d2 = d[k] * d[k];
            const float diff = d2 - ss;
            worst = fmaxf(worst, fabsf(diff) / d2);
            const float sr = s[0] * r0[k][0] + s[1] * r0[k][1] + s[2] * r0[k][2];
            const float g = diff / (2.f * sr * (inv_m0 + inv_m[k]));
            for (int x = 0; x < 3; ++x) {
                p[k][x] += g * r0[k][x] * inv_m[k];
                p0[x] -= g * r0[k][x] * inv_m0;
            }
        }
        if (worst < tol) return it;
    }
    return max_iter + 1;
}
