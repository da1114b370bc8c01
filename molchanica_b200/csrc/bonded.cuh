// bonded.cuh -- host-side launcher of bonded.cu
#pragma once
#include "common.cuh"

struct BondedTerms {
    int n_bonds = 0, n_angles = 0, n_dihedrals = 0;
    const int2 *bonds = nullptr;         // (i, j) original ids
    const float2 *bond_kr0 = nullptr;    // k [kcal/mol/A^2], r0 [A]
    const int4 *angles = nullptr;        // (i, j, k, -) with j the vertex
    const float2 *angle_kt0 = nullptr;   // k [kcal/mol/rad^2], theta0 [rad]
    const int4 *dihedrals = nullptr;     // (i, j, k, l)
    const float4 *dihedral_prm = nullptr;  // pk [kcal/mol], periodicity, phase [rad], -
    // Decomposed rank: every rank holds the whole term list and evaluates the terms that touch an atom it owns (slots
    // own0 .. own1-1); partners are owned or ghosts (a term spans a few Angstrom, a ghost layer is >= cutoff + skin deep).
    // Forces go to owned atoms only -- the neighbour evaluates the same term for its own atoms, as with the full pair lists:
    // no reverse communication -- and the term's energy / virial is shared out by the fraction of its atoms owned here, so
    // that the sum over ranks counts it once.  A term with an owned atom and a partner this rank does not hold raises *missing.
    int own0 = 0, own1 = 0x7fffffff;
    int *missing = nullptr;
};

// Adds the bonded forces to `force` (cell-order slots) and, when want_energy, writes {E_bond, E_angle, E_dihedral}
// to energy3 (device, 3 doubles).
void launch_bonded(const BondedTerms &t, const int *slot_of_orig, const float4 *xyzq, const NbParams &p, float4 *force,
                   double *energy3, bool want_energy, cudaStream_t st, int64_t *launches);
