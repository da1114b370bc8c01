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
