// group_energy.cu -- interaction energy between molecules on demand (SnapshotEnergyData.energy_potential_between_mols,
// reference src/md/mod.rs:1242-1245): the sum of the nonbonded pair energies over the listed pairs whose atoms carry
// different molecule ids.  One thread per list row, the row's partners gathered like the force kernel does; the
// cutoff decision uses the same fp32 r^2 expression.  Runs only when mc_get_energy_between_mols is called.
// STATUS: arithmetic and kernel source verified on the host (tests/test_kernels_on_host.py), not yet run on hardware.
#include "group_energy.cuh"
#include "pair_energy_terms.h"

namespace {

__global__ void __launch_bounds__(128) between_mols_kernel(int n_rows, int row0, const float4 *__restrict__ xyzq,
                                                            const uint16_t *__restrict__ type, const int *__restrict__ orig,
                                                            const uint16_t *__restrict__ mol_of_orig,
                                                            const uint32_t *__restrict__ nbr_start, const uint32_t *__restrict__ nbr_count,
                                                            const uint32_t *__restrict__ nbr_list, const float2 *__restrict__ ljtab,
                                                            const NbParams p, int lj_on, int coul_mode, double *__restrict__ energy) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    float e = 0.f;
    if (r < n_rows) {
        const int i = row0 + r;
        const float4 xi = xyzq[i];
        const int ti = type[i];
        const uint16_t mi = mol_of_orig[orig[i]];
        const uint32_t s = nbr_start[i], cnt = nbr_count[i];
        for (uint32_t k = 0; k < cnt; ++k) {
            const uint32_t j = nbr_list[s + k];
            if (mol_of_orig[orig[j]] == mi) continue;
            const float4 xj = xyzq[j];
            float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
            if (p.periodic) {
                dx = __fmaf_rn(-rintf(dx * p.inv_ext[0]), p.ext[0], dx);
                dy = __fmaf_rn(-rintf(dy * p.inv_ext[1]), p.ext[1], dy);
                dz = __fmaf_rn(-rintf(dz * p.inv_ext[2]), p.ext[2], dz);
            }
            const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const float2 lj = ljtab[ti * p.n_types + type[j]];
            e += mc_pair_energy(r2, lj.x, lj.y, xi.w * xj.w, p.rc2_lj, p.rc2_q, lj_on, coul_mode, p.alpha);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) e += __shfl_xor_sync(MC_FULL_MASK, e, d);
    if ((threadIdx.x & 31) == 0 && e != 0.f) atomicAdd(energy, 0.5 * (double)e);  // every pair sits in two rows
}

// Virial of the nonbonded terms on demand (mc_get_pressure): W = sum_{i<j} r_ij . f_ij over the listed pairs inside the
// cutoffs plus the scaled 1-4 pairs, with the forms of pair_force.cu (same fp32 r^2 expression for the cutoff decision).
// Eight lanes per row like the force kernel (consecutive lanes read consecutive list entries); every pair sits in two
// rows, hence the factor 1/2.
constexpr int VIR_LANES = 8;
__global__ void __launch_bounds__(128) virial_kernel(int n_rows, int row0, const float4 *__restrict__ xyzq, const uint16_t *__restrict__ type,
                                                      const int *__restrict__ orig, const int *__restrict__ slot_of_orig,
                                                      const uint32_t *__restrict__ nbr_start, const uint32_t *__restrict__ nbr_count,
                                                      const uint32_t *__restrict__ nbr_list, const int32_t *__restrict__ p14_start,
                                                      const int32_t *__restrict__ p14_idx, const float2 *__restrict__ ljtab, const NbParams p,
                                                      int lj_on, int coul_mode, float scale_lj, float scale_q, double *__restrict__ virial) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = gt / VIR_LANES, lane = gt % VIR_LANES;
    float w = 0.f;
    if (r < n_rows) {
        const int i = row0 + r;
        const float4 xi = xyzq[i];
        const int ti = type[i];
        const uint32_t s = nbr_start[i], cnt = nbr_count[i];
        for (uint32_t k = lane; k < cnt; k += VIR_LANES) {
            const uint32_t j = nbr_list[s + k];
            const float4 xj = xyzq[j];
            float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
            if (p.periodic) {
                dx = __fmaf_rn(-rintf(dx * p.inv_ext[0]), p.ext[0], dx);
                dy = __fmaf_rn(-rintf(dy * p.inv_ext[1]), p.ext[1], dy);
                dz = __fmaf_rn(-rintf(dz * p.inv_ext[2]), p.ext[2], dz);
            }
            const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const float2 lj = ljtab[ti * p.n_types + type[j]];
            w += mc_pair_virial(r2, lj.x, lj.y, xi.w * xj.w, p.rc2_lj, p.rc2_q, lj_on, coul_mode, p.alpha);
        }
        if (p14_start && lane == 0) {  // Amber 1-4 rows: no cutoff, LJ x scale_lj, plain Coulomb x scale_q (pairs14_kernel)
            const int oi = orig[i];
            for (int e = p14_start[oi]; e < p14_start[oi + 1]; ++e) {
                const int j = slot_of_orig[p14_idx[e]];
                if (j < 0) continue;
                const float4 xj = xyzq[j];
                float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                if (p.periodic) {
                    dx -= rintf(dx * p.inv_ext[0]) * p.ext[0];
                    dy -= rintf(dy * p.inv_ext[1]) * p.ext[1];
                    dz -= rintf(dz * p.inv_ext[2]) * p.ext[2];
                }
                const float r2 = dx * dx + dy * dy + dz * dz;
                const float2 lj = ljtab[ti * p.n_types + type[j]];
                const float big = 3.0e38f;
                w += scale_lj * mc_pair_virial(r2, lj.x, lj.y, 0.f, big, big, lj_on, 0, 0.f) +
                     scale_q * mc_pair_virial(r2, 1.f, 0.f, xi.w * xj.w, big, big, 0, coul_mode != 0 ? 1 : 0, 0.f);
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(MC_FULL_MASK, w, d);
    if ((threadIdx.x & 31) == 0 && w != 0.f) atomicAdd(virial, 0.5 * (double)w);
}

}  // namespace

#ifdef MC_HAVE_LAUNCH  // the stand-ins of tests/cpp/shim/ and shim_mt/ have no launcher; shim_fiber/ has
void launch_between_mols(int n_rows, int row0, const float4 *xyzq, const uint16_t *type, const int *orig, const uint16_t *mol_of_orig,
                         const uint32_t *nbr_start, const uint32_t *nbr_count, const uint32_t *nbr_list, const float2 *ljtab,
                         const NbParams &p, int lj_on, int coul_mode, double *energy, cudaStream_t st, int64_t *launches) {
    cudaMemsetAsync(energy, 0, sizeof(double), st);
    if (n_rows <= 0) return;
    MC_LAUNCH(between_mols_kernel, div_up((size_t)n_rows, 128), 128, 0, st, n_rows, row0, xyzq, type, orig, mol_of_orig, nbr_start, nbr_count,
                                                                     nbr_list, ljtab, p, lj_on, coul_mode, energy);
    *launches += 1;
}

void launch_virial(int n_rows, int row0, const float4 *xyzq, const uint16_t *type, const int *orig, const int *slot_of_orig,
                   const uint32_t *nbr_start, const uint32_t *nbr_count, const uint32_t *nbr_list, const int32_t *p14_start,
                   const int32_t *p14_idx, const float2 *ljtab, const NbParams &p, int lj_on, int coul_mode, float scale_lj, float scale_q,
                   double *virial, cudaStream_t st, int64_t *launches) {
    cudaMemsetAsync(virial, 0, sizeof(double), st);
    if (n_rows <= 0) return;
    MC_LAUNCH(virial_kernel, div_up((size_t)n_rows * VIR_LANES, 128), 128, 0, st, n_rows, row0, xyzq, type, orig, slot_of_orig, nbr_start, nbr_count,
              nbr_list, p14_start, p14_idx, ljtab, p, lj_on, coul_mode, scale_lj, scale_q, virial);
    *launches += 1;
}
#endif
