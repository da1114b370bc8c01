// Test infrastructure: compiles molchanica_b200/csrc/bonded_terms.h -- the very arithmetic bonded.cu runs on the
// GPU -- with g++ and exposes it to the Python tests, which compare it with the fp64 oracle and with finite
// differences on a machine without a GPU.  Not part of the product library (there is no CPU path in it).
#include <stdint.h>
#include <string.h>

#include "../../molchanica_b200/csrc/bonded_terms.h"

extern "C" {

// forces: n x 3 floats (accumulated), energy3: {bond, angle, dihedral}; ids index xyz (n x 3 floats), no PBC
void bonded_host_eval(const float *xyz, int64_t nb, const int32_t *bonds, const float *kr0, int64_t na, const int32_t *angles,
                      const float *kt0, int64_t nd, const int32_t *dih, const float *prm, float *forces, double *energy3) {
    energy3[0] = energy3[1] = energy3[2] = 0.0;
    auto sub = [&](int i, int j, float d[3]) { for (int a = 0; a < 3; ++a) d[a] = xyz[3 * i + a] - xyz[3 * j + a]; };
    auto add = [&](int i, const float f[3], float s) { for (int a = 0; a < 3; ++a) forces[3 * i + a] += s * f[a]; };
    for (int64_t t = 0; t < nb; ++t) {
        const int i = bonds[2 * t], j = bonds[2 * t + 1];
        float d[3], fi[3];
        sub(i, j, d);
        energy3[0] += mc_bond_term(d, kr0[2 * t], kr0[2 * t + 1], fi);
        add(i, fi, 1.f); add(j, fi, -1.f);
    }
    for (int64_t t = 0; t < na; ++t) {
        const int i = angles[3 * t], j = angles[3 * t + 1], k = angles[3 * t + 2];
        float a[3], b[3], fi[3], fk[3];
        sub(i, j, a); sub(k, j, b);
        energy3[1] += mc_angle_term(a, b, kt0[2 * t], kt0[2 * t + 1], fi, fk);
        add(i, fi, 1.f); add(k, fk, 1.f); add(j, fi, -1.f); add(j, fk, -1.f);
    }
    for (int64_t t = 0; t < nd; ++t) {
        const int i = dih[4 * t], j = dih[4 * t + 1], k = dih[4 * t + 2], l = dih[4 * t + 3];
        float rij[3], rkj[3], rkl[3], fi[3], fj[3], fk[3], fl[3];
        sub(i, j, rij); sub(k, j, rkj); sub(k, l, rkl);
        energy3[2] += mc_dihedral_term(rij, rkj, rkl, prm[3 * t], prm[3 * t + 1], prm[3 * t + 2], fi, fj, fk, fl);
        add(i, fi, 1.f); add(j, fj, 1.f); add(k, fk, 1.f); add(l, fl, 1.f);
    }
}
}
