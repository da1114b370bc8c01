// Test infrastructure: compiles molchanica_b200/csrc/pme_terms.h -- the arithmetic pme.cu runs on the GPU -- with g++
// and mirrors the kernels' loops (spread, influence, gather, exclusion correction) on host arrays for
// tests/test_pme_cpu.py; the FFT in between is numpy's.  Not part of the product library.
#include <stdint.h>
#include <string.h>

#include "../../molchanica_b200/csrc/pme_terms.h"

extern "C" {

void pme_host_spread(int64_t n, const float *xyzq, const float *lo, const float *ext, const int *K, float *grid) {
    for (int64_t i = 0; i < n; ++i) {
        const float *p = xyzq + 4 * i;
        if (p[3] == 0.f) continue;
        int k0[3];
        float th[3][4], dth[4], w;
        for (int a = 0; a < 3; ++a) { mc_pme_coord(p[a], lo[a], 1.0f / ext[a], K[a], &k0[a], &w); mc_bspline4(w, th[a], dth); }
        for (int a = 0; a < 4; ++a) {
            const int ia = mc_pme_wrap(k0[0], a, K[0]);
            const float qa = p[3] * th[0][a];
            for (int b = 0; b < 4; ++b) {
                const int ib = mc_pme_wrap(k0[1], b, K[1]);
                const float qab = qa * th[1][b];
                float *row = grid + ((size_t)ia * K[1] + ib) * K[2];
                for (int c = 0; c < 4; ++c) row[mc_pme_wrap(k0[2], c, K[2])] += qab * th[2][c];
            }
        }
    }
}

void pme_host_influence(const int *K, const float *ext, float alpha, float *bc /* K1 x K2 x (K3/2+1) */) {
    const int K3h = K[2] / 2 + 1;
    const float inv_ext[3] = {1.0f / ext[0], 1.0f / ext[1], 1.0f / ext[2]};
    const double vol = (double)ext[0] * ext[1] * ext[2];
    const float pi = 3.14159265358979f;
    for (int i1 = 0; i1 < K[0]; ++i1)
        for (int i2 = 0; i2 < K[1]; ++i2)
            for (int i3 = 0; i3 < K3h; ++i3)
                bc[((size_t)i1 * K[1] + i2) * K3h + i3] =
                    mc_pme_influence(i1, i2, i3, K[0], K[1], K[2], inv_ext, (float)(1.0 / (3.14159265358979323846 * vol)),
                                     pi * pi / (alpha * alpha), (float)mc_pme_bmod4(i1, K[0]), (float)mc_pme_bmod4(i2, K[1]),
                                     (float)mc_pme_bmod4(i3, K[2]));
}

void pme_host_gather(int64_t n, const float *xyzq, const float *lo, const float *ext, const int *K, const float *grid, float *force) {
    for (int64_t i = 0; i < n; ++i) {
        const float *p = xyzq + 4 * i;
        if (p[3] == 0.f) continue;
        int k0[3];
        float th[3][4], dth[3][4], w;
        for (int a = 0; a < 3; ++a) { mc_pme_coord(p[a], lo[a], 1.0f / ext[a], K[a], &k0[a], &w); mc_bspline4(w, th[a], dth[a]); }
        float fx = 0.f, fy = 0.f, fz = 0.f;
        for (int a = 0; a < 4; ++a) {
            const int ia = mc_pme_wrap(k0[0], a, K[0]);
            for (int b = 0; b < 4; ++b) {
                const int ib = mc_pme_wrap(k0[1], b, K[1]);
                const float *row = grid + ((size_t)ia * K[1] + ib) * K[2];
                for (int c = 0; c < 4; ++c) {
                    const float phi = row[mc_pme_wrap(k0[2], c, K[2])];
                    fx += phi * dth[0][a] * th[1][b] * th[2][c];
                    fy += phi * th[0][a] * dth[1][b] * th[2][c];
                    fz += phi * th[0][a] * th[1][b] * dth[2][c];
                }
            }
        }
        force[3 * i] -= p[3] * fx * ((float)K[0] / ext[0]);
        force[3 * i + 1] -= p[3] * fy * ((float)K[1] / ext[1]);
        force[3 * i + 2] -= p[3] * fz * ((float)K[2] / ext[2]);
    }
}

double pme_host_excl(int64_t n, const float *xyzq, const float *ext, int periodic, const int32_t *excl_start, const int32_t *excl_idx,
                     float alpha, float *force) {
    double e = 0.0;
    for (int64_t i = 0; i < n; ++i)
        for (int t = excl_start[i]; t < excl_start[i + 1]; ++t) {
            const int j = excl_idx[t];
            if (j == i) continue;
            float d[3], f[3];
            for (int a = 0; a < 3; ++a) {
                d[a] = xyzq[4 * i + a] - xyzq[4 * j + a];
                if (periodic) d[a] -= rintf(d[a] * (1.0f / ext[a])) * ext[a];
            }
            e += 0.5 * (double)mc_pme_excl_term(d, xyzq[4 * i + 3] * xyzq[4 * j + 3], alpha, f);
            for (int a = 0; a < 3; ++a) force[3 * i + a] += f[a];
        }
    return e;
}
}
