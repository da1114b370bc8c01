// group_energy.cuh -- host-side launcher of group_energy.cu
#pragma once
#include "common.cuh"

// *energy (device double) = sum over listed pairs with different molecule ids of the nonbonded pair energy
void launch_between_mols(int n_rows, int row0, const float4 *xyzq, const uint16_t *type, const int *orig, const uint16_t *mol_of_orig,
                         const uint32_t *nbr_start, const uint32_t *nbr_count, const uint32_t *nbr_list, const float2 *ljtab,
                         const NbParams &p, int lj_on, int coul_mode, double *energy, cudaStream_t st, int64_t *launches);

// *virial (device double) = sum_{i<j} r_ij . f_ij over the listed pairs inside the cutoffs + the scaled 1-4 pairs (p14_start may be null)
void launch_virial(int n_rows, int row0, const float4 *xyzq, const uint16_t *type, const int *orig, const int *slot_of_orig,
                   const uint32_t *nbr_start, const uint32_t *nbr_count, const uint32_t *nbr_list, const int32_t *p14_start,
                   const int32_t *p14_idx, const float2 *ljtab, const NbParams &p, int lj_on, int coul_mode, float scale_lj, float scale_q,
                   double *virial, cudaStream_t st, int64_t *launches);
