// Test infrastructure: compiles molchanica_b200/csrc/settle_terms.h -- the arithmetic settle.cu runs on the GPU --
// with g++ for the CPU tests (tests/test_settle_cpu.py).  Not part of the product library.
#include <stdint.h>

#include "../../molchanica_b200/csrc/settle_terms.h"

extern "C" {
// x0: n x 9 old positions (O, H1, H2), x1: n x 9 unconstrained new positions, out: n x 9 constrained positions
void settle_host_eval(int64_t n, const float *x0, const float *x1, float m_o, float m_h, float d_oh, float d_hh, float *out) {
    const SettleParams p = mc_settle_params(m_o, m_h, d_oh, d_hh);
    for (int64_t w = 0; w < n; ++w) {
        const float *o = x0 + 9 * w, *q = x1 + 9 * w;
        float b0[3], c0[3], a1[3], b1[3], c1[3], a3[3], b3[3], c3[3];
        for (int a = 0; a < 3; ++a) {
            b0[a] = o[3 + a] - o[a]; c0[a] = o[6 + a] - o[a];
            a1[a] = q[a] - o[a]; b1[a] = q[3 + a] - o[a]; c1[a] = q[6 + a] - o[a];
        }
        mc_settle(p, b0, c0, a1, b1, c1, a3, b3, c3);
        for (int a = 0; a < 3; ++a) {
            out[9 * w + a] = o[a] + a3[a]; out[9 * w + 3 + a] = o[a] + b3[a]; out[9 * w + 6 + a] = o[a] + c3[a];
        }
    }
}
}

// The body of settle_kernel (molchanica_b200/csrc/settle.cu), line for line, on host arrays: x1 = positions after the
// unconstrained drift (n x 9), v = velocities (n x 9), ext = periodic box (or NULL).  Updates x1 and v in place.
extern "C" void settle_host_step(int64_t n, float *x1, float *v, const float *ext, float m_o, float m_h, float d_oh, float d_hh,
                                 float dt) {
    const SettleParams sp = mc_settle_params(m_o, m_h, d_oh, d_hh);
    for (int64_t w = 0; w < n; ++w) {
        float *xo_ = x1 + 9 * w, *x1_ = xo_ + 3, *x2_ = xo_ + 6;
        float *vo_ = v + 9 * w, *v1_ = vo_ + 3, *v2_ = vo_ + 6;
        float b0[3], c0[3], a1[3], b1[3], c1[3], a3[3], b3[3], c3[3];
        for (int a = 0; a < 3; ++a) {
            float db = x1_[a] - xo_[a], dc = x2_[a] - xo_[a];
            if (ext) {
                const float inv = 1.f / ext[a];
                db -= rintf(db * inv) * ext[a];
                dc -= rintf(dc * inv) * ext[a];
            }
            a1[a] = vo_[a] * dt;
            b0[a] = db - (v1_[a] - vo_[a]) * dt;
            c0[a] = dc - (v2_[a] - vo_[a]) * dt;
            b1[a] = b0[a] + v1_[a] * dt;
            c1[a] = c0[a] + v2_[a] * dt;
        }
        mc_settle(sp, b0, c0, a1, b1, c1, a3, b3, c3);
        const float inv_dt = 1.f / dt;
        for (int a = 0; a < 3; ++a) {
            const float da = a3[a] - a1[a], db = b3[a] - b1[a], dc = c3[a] - c1[a];
            xo_[a] += da; x1_[a] += db; x2_[a] += dc;
            vo_[a] += da * inv_dt; v1_[a] += db * inv_dt; v2_[a] += dc * inv_dt;
        }
    }
}

#include "../../molchanica_b200/csrc/vsite_terms.h"
// virtual-site arithmetic on host arrays: x n x 12 (O, H1, H2, M), f n x 12
extern "C" void vsite_host_eval(int64_t n, float *x, float *f, float a, float b) {
    for (int64_t w = 0; w < n; ++w) {
        float *o = x + 12 * w, d1[3], d2[3], m[3], fo[3], f1[3], f2[3];
        for (int k = 0; k < 3; ++k) { d1[k] = o[3 + k] - o[k]; d2[k] = o[6 + k] - o[k]; }
        mc_vsite_position(o, d1, d2, a, b, m);
        for (int k = 0; k < 3; ++k) o[9 + k] = m[k];
        float *ff = f + 12 * w;
        mc_vsite_spread(ff + 9, a, b, fo, f1, f2);
        for (int k = 0; k < 3; ++k) { ff[k] += fo[k]; ff[3 + k] += f1[k]; ff[6 + k] += f2[k]; ff[9 + k] = 0.f; }
    }
}

#include "../../molchanica_b200/csrc/shake_terms.h"
// SHAKE arithmetic of one cluster per row: x0 / x1 n x 12 (heavy, h1, h2, h3; unused hydrogens ignored via nh[i]), out n x 12
extern "C" int shake_host_eval(int64_t n, const int32_t *nh, const float *x0, const float *x1, const float *inv_m /* n x 4 */,
                               const float *d /* n x 3 */, float tol, float *out) {
    int worst = 0;
    for (int64_t c = 0; c < n; ++c) {
        const float *o = x0 + 12 * c, *q = x1 + 12 * c;
        float r0[3][3], p[3][3], p0[3], im[3], dd[3];
        for (int a = 0; a < 3; ++a) p0[a] = q[a] - o[a];
        for (int k = 0; k < nh[c]; ++k) {
            im[k] = inv_m[4 * c + 1 + k]; dd[k] = d[3 * c + k];
            for (int a = 0; a < 3; ++a) { r0[k][a] = o[3 + 3 * k + a] - o[a]; p[k][a] = q[3 + 3 * k + a] - o[a]; }
        }
        const int it = mc_shake_cluster(nh[c], r0, p0, p, inv_m[4 * c], im, dd, tol, 64);
        if (it > worst) worst = it;
        for (int a = 0; a < 3; ++a) out[12 * c + a] = o[a] + p0[a];
        for (int k = 0; k < 3; ++k)
            for (int a = 0; a < 3; ++a) out[12 * c + 3 + 3 * k + a] = k < nh[c] ? o[a] + p[k][a] : q[3 + 3 * k + a];
    }
    return worst;
}
