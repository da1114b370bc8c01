// pme.cuh -- host-side interface of pme.cu (SPME reciprocal space)
#pragma once
#include "common.cuh"

struct PmeState {
    int K[3] = {0, 0, 0};
    int plan_r2c = 0, plan_c2r = 0;
    bool planned = false;
    float *grid = nullptr;        // K1 x K2 x K3 real charge grid / convolved potential
    float2 *cgrid = nullptr;      // K1 x K2 x (K3/2+1)
    float *bmod[3] = {nullptr, nullptr, nullptr};
    double *energy = nullptr;     // device: {E_recip, E_exclusion_correction, W_recip, W_exclusion_correction} (W: virial)
    double self_q2 = 0.0;         // sum q_i^2 (host)
};

// Returns 0 or a negative MC_E_* code; msg explains (cuFFT missing, plan failure ...).
int pme_configure(PmeState *s, int k1, int k2, int k3, cudaStream_t st, const char **msg);
void pme_release(PmeState *s);
// Adds the reciprocal-space forces to `force` (cell-order slots, rows row0 .. row0 + n_rows - 1 = all atoms on a
// single GPU) and, when want_energy, writes E_recip to s->energy[0].
int pme_launch(PmeState *s, int n, const float4 *xyzq, const float lo[3], const float ext[3], float alpha, float4 *force,
               bool want_energy, cudaStream_t st, int64_t *launches, const char **msg);
// erf correction of the excluded pairs (CSR in original ids, every pair in both rows): adds forces, energy to s->energy[1]
void pme_launch_exclusions(PmeState *s, int n, const float4 *xyzq, const int *orig, const int *slot_of_orig,
                           const int32_t *excl_start, const int32_t *excl_idx, const NbParams &p, float4 *force,
                           bool want_energy, cudaStream_t st, int64_t *launches);
