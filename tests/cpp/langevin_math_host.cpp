// Test infrastructure: compiles molchanica_b200/csrc/langevin_terms.h -- the generator and the O step the GPU kernel
// runs -- with g++ for tests/test_langevin_cpu.py.  Not part of the product library.
#include "../../molchanica_b200/csrc/langevin_terms.h"

extern "C" {
void lgv_host_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { mc_philox4x32_10(ctr, key, out); }
void lgv_host_normals(uint64_t seed, int64_t n, uint64_t step, float *xi) {
    for (int64_t i = 0; i < n; ++i) mc_langevin_normals(seed, (uint32_t)i, step, xi + 3 * i);
}
// the body of langevin_ou_kernel on host arrays: vel n x 4 (vx, vy, vz, 1/m), ids n
void lgv_host_ou(int64_t n, float *vel, const int32_t *ids, float c1, float c2, float kT, uint64_t seed, uint64_t step) {
    for (int64_t i = 0; i < n; ++i) {
        float *v = vel + 4 * i;
        if (v[3] <= 0.f) continue;
        float xi[3], vv[3] = {v[0], v[1], v[2]};
        mc_langevin_normals(seed, (uint32_t)ids[i], step, xi);
        mc_langevin_ou(vv, v[3], c1, c2, kT, xi);
        v[0] = vv[0]; v[1] = vv[1]; v[2] = vv[2];
    }
}
}

#include "../../molchanica_b200/csrc/csvr_terms.h"
extern "C" double csvr_host_lambda(double kinetic, double kT, double nf, double c, uint64_t seed, uint64_t step) {
    return mc_csvr_lambda(kinetic, kT, nf, c, seed, step);
}
