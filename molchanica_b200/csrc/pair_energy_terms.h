// pair_energy_terms.h -- energy of one nonbonded pair, the forms of pair_force.cu restated for the on-demand energy
// decompositions of the snapshot (SnapshotEnergyData.energy_potential_between_mols, reference src/md/mod.rs:1242-1245,
// ui/panels/md_viewer.rs:202-256).  Not used by the step path.  Shared by device (group_energy.cu) and host tests.
//   LJ       4 eps ((sigma/r)^12 - (sigma/r)^6)            for r^2 < rc_lj^2     (table holds sigma^2 and 24 eps)
//   Coulomb  plain: q_i q_j / r;  Ewald real space: q_i q_j erfc(alpha r) / r   for r^2 < rc_q^2
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MC_PE_HD __host__ __device__ __forceinline__
#else
#define MC_PE_HD inline
#endif

// coul_mode: 0 none, 1 plain, 2 erfc (MC_COULOMB_*).  lj_on = 0 drops the LJ part.
MC_PE_HD float mc_pair_energy(float r2, float sig2, float eps24, float qq, float rc2_lj, float rc2_q, int lj_on, int coul_mode,
                              float alpha) {
    float e = 0.f;
    if (lj_on && r2 < rc2_lj) {
        const float s2 = sig2 / r2, s6 = s2 * s2 * s2;
        e += eps24 * (1.f / 6.f) * s6 * (s6 - 1.f);
    }
    if (coul_mode != 0 && r2 < rc2_q) {
        const float r = sqrtf(r2);
        e += coul_mode == 1 ? qq / r : qq * erfcf(alpha * r) / r;
    }
    return e;
}

// r_ij . f_ij of the same pair (the pair's contribution to the virial W = sum_{i<j} r_ij . f_ij, pressure =
// (2 KE + W) / 3V; SnapshotEnergyData.pressure, reference ui/panels/md_viewer.rs:202-256):
//   LJ       24 eps (2 (sigma/r)^12 - (sigma/r)^6)
//   Coulomb  plain: q_i q_j / r;  Ewald real space: q_i q_j (erfc(alpha r)/r + 2 alpha/sqrt(pi) exp(-alpha^2 r^2))
MC_PE_HD float mc_pair_virial(float r2, float sig2, float eps24, float qq, float rc2_lj, float rc2_q, int lj_on, int coul_mode,
                              float alpha) {
    float w = 0.f;
    if (lj_on && r2 < rc2_lj) {
        const float s2 = sig2 / r2, s6 = s2 * s2 * s2;
        w += eps24 * s6 * (2.f * s6 - 1.f);
    }
    if (coul_mode != 0 && r2 < rc2_q) {
        const float r = sqrtf(r2);
        w += coul_mode == 1 ? qq / r : qq * (erfcf(alpha * r) / r + 1.1283791670955126f * alpha * expf(-alpha * alpha * r2));
    }
    return w;
}
