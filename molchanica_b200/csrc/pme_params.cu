// pme_params.cu -- host-only helper: Ewald splitting parameter and SPME grid for a cutoff and an error tolerance
// (the choice the reference delegates to its `ewald` crate, Cargo.toml:30).  No device code; usable without a GPU.
//   alpha: smallest value with erfc(alpha rc) / rc <= tol (the relative size of the neglected real-space tail)
//   K_a:   >= 2 alpha L_a / (3 tol^(1/5))  (the OpenMM / Essmann rule of thumb for order-4..5 splines),
//          rounded up to a product of 2, 3, 5 and 7 so that the FFT stays fast, at least 8
#include <cmath>
#include <initializer_list>

#include "../../include/molchanica_md.h"

static int next_smooth(int n) {
    for (int k = n < 8 ? 8 : n;; ++k) {
        int m = k;
        for (int p : {2, 3, 5, 7})
            while (m % p == 0) m /= p;
        if (m == 1) return k;
    }
}

extern "C" int mc_pme_suggest(float rc, float tol, const float box_ext[3], float *alpha, int32_t grid[3]) {
    if (!(rc > 0.f) || !(tol > 0.f && tol < 1.f) || !box_ext || !alpha || !grid) return MC_E_INVALID;
    double lo = 0.0, hi = 1.0;
    while (std::erfc(hi * rc) / rc > tol) hi *= 2.0;
    for (int it = 0; it < 200; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (std::erfc(mid * rc) / rc > tol) lo = mid; else hi = mid;
    }
    *alpha = (float)hi;
    for (int a = 0; a < 3; ++a) {
        if (!(box_ext[a] > 0.f)) return MC_E_INVALID;
        const double k = 2.0 * hi * (double)box_ext[a] / (3.0 * std::pow((double)tol, 0.2));
        grid[a] = next_smooth((int)std::ceil(k));
    }
    return MC_OK;
}
