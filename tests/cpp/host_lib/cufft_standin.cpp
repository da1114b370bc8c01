// Test infrastructure: the five cuFFT entry points pme.cu resolves with dlopen, as a plain separable DFT on host memory
// (O(K^4), exact to double rounding: grids of 16-32 points per axis).  Same layouts and sign conventions as cuFFT:
// R2C forward with exp(-2 pi i ..), out[k1][k2][k3/2+1]; C2R inverse, unnormalised, reading the half spectrum.
// MOLCHANICA_CUFFT_LIB points the host build of the library at this file's shared object.
#include <math.h>
#include <stdlib.h>

#include <complex>
#include <map>
#include <vector>

namespace {
struct Plan { int k1, k2, k3, type; };
std::map<int, Plan> plans;
int next_id = 1;
typedef std::complex<double> cd;

void dft_axis(std::vector<cd> &a, int n0, int n1, int n2, int axis, int sign) {
    const int n[3] = {n0, n1, n2};
    const int len = n[axis];
    std::vector<cd> tw((size_t)len), line((size_t)len), out((size_t)len);
    for (int k = 0; k < len; ++k) tw[(size_t)k] = std::polar(1.0, sign * 2.0 * M_PI * k / len);
    const size_t stride = axis == 0 ? (size_t)n1 * n2 : axis == 1 ? (size_t)n2 : 1;
    const int o1 = axis == 0 ? n1 : n0, o2 = axis == 2 ? n1 : n2;
    for (int i = 0; i < o1; ++i)
        for (int j = 0; j < o2; ++j) {
            size_t base;
            if (axis == 0) base = (size_t)i * n2 + j;
            else if (axis == 1) base = (size_t)i * n1 * n2 + j;
            else base = ((size_t)i * n1 + j) * n2;
            for (int k = 0; k < len; ++k) line[(size_t)k] = a[base + (size_t)k * stride];
            for (int m = 0; m < len; ++m) {
                cd s = 0;
                for (int k = 0; k < len; ++k) s += line[(size_t)k] * tw[(size_t)(((long long)m * k) % len)];
                out[(size_t)m] = s;
            }
            for (int k = 0; k < len; ++k) a[base + (size_t)k * stride] = out[(size_t)k];
        }
}
}  // namespace

extern "C" {
int cufftPlan3d(int *plan, int k1, int k2, int k3, int type) {
    *plan = next_id++;
    plans[*plan] = Plan{k1, k2, k3, type};
    return 0;
}
int cufftSetStream(int, void *) { return 0; }
int cufftDestroy(int plan) { plans.erase(plan); return 0; }

int cufftExecR2C(int plan, float *in, float *out /* float2 */) {
    auto it = plans.find(plan);
    if (it == plans.end()) return 1;
    const Plan p = it->second;
    const size_t n = (size_t)p.k1 * p.k2 * p.k3;
    std::vector<cd> a(n);
    for (size_t i = 0; i < n; ++i) a[i] = in[i];
    dft_axis(a, p.k1, p.k2, p.k3, 2, -1);
    dft_axis(a, p.k1, p.k2, p.k3, 1, -1);
    dft_axis(a, p.k1, p.k2, p.k3, 0, -1);
    const int h = p.k3 / 2 + 1;
    for (int i = 0; i < p.k1; ++i)
        for (int j = 0; j < p.k2; ++j)
            for (int k = 0; k < h; ++k) {
                const cd v = a[((size_t)i * p.k2 + j) * p.k3 + k];
                float *o = out + 2 * (((size_t)i * p.k2 + j) * h + k);
                o[0] = (float)v.real();
                o[1] = (float)v.imag();
            }
    return 0;
}

int cufftExecC2R(int plan, float *in /* float2 */, float *out) {
    auto it = plans.find(plan);
    if (it == plans.end()) return 1;
    const Plan p = it->second;
    const size_t n = (size_t)p.k1 * p.k2 * p.k3;
    const int h = p.k3 / 2 + 1;
    std::vector<cd> a(n);
    // full spectrum from the half: X[-k] = conj(X[k])
    for (int i = 0; i < p.k1; ++i)
        for (int j = 0; j < p.k2; ++j)
            for (int k = 0; k < p.k3; ++k) {
                cd v;
                if (k < h) {
                    const float *s = in + 2 * (((size_t)i * p.k2 + j) * h + k);
                    v = cd(s[0], s[1]);
                } else {
                    const int ii = (p.k1 - i) % p.k1, jj = (p.k2 - j) % p.k2, kk = p.k3 - k;
                    const float *s = in + 2 * (((size_t)ii * p.k2 + jj) * h + kk);
                    v = cd(s[0], -s[1]);
                }
                a[((size_t)i * p.k2 + j) * p.k3 + k] = v;
            }
    dft_axis(a, p.k1, p.k2, p.k3, 2, +1);
    dft_axis(a, p.k1, p.k2, p.k3, 1, +1);
    dft_axis(a, p.k1, p.k2, p.k3, 0, +1);
    for (size_t i = 0; i < n; ++i) out[i] = (float)a[i].real();
    return 0;
}
}
