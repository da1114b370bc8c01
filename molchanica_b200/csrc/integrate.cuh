// integrate.cuh -- host-side launchers of integrate.cu
#pragma once
#include "common.cuh"
#include "halo_sync.cuh"

// v += (F + F_ext) / m * kick * 418.4 ; then (drift != 0) x += v * drift and the displacement
// since the last list build is checked against max_disp minus a look-ahead margin of `lookahead`
// further drifts (flag raised when exceeded).  host_flag != nullptr: the last block to finish writes
// (step_tag << 2 | flag bits) into that pinned host word (done_counter: one zeroed device word).
void launch_kick_drift(int n_rows, float4 *xyzq, float4 *vel, const float4 *force, const float *ext_force,
                       const int *orig, const uint8_t *flags, const float4 *xref, float kick, float drift,
                       float max_disp, float lookahead, int *rebuild_flag, cudaStream_t st, int64_t *launches,
                       uint32_t *done_counter = nullptr, int *host_flag = nullptr, int step_tag = 0);
// The decomposed step: the same kick + drift, with the boundary layers' new positions also stored into the
// neighbour ranks' ghost blocks over NVLink and the ack / ready flags of halo_sync.cuh raised in-kernel.
void launch_kick_drift_halo(int n_rows, float4 *xyzq, float4 *vel, const float4 *force, const float *ext_force,
                            const int *orig, const uint8_t *flags, const float4 *xref, float kick, float drift,
                            float max_disp, int *rebuild_flag, const HaloPush &hp, cudaStream_t st, int64_t *launches);
// Langevin O step: v <- c1 v + c2 sqrt(kT/m) xi with Philox noise keyed by (seed, original atom id, step)
void launch_langevin_ou(int n_rows, float4 *vel, const int *orig, const uint8_t *flags, float c1, float c2, float kT, uint64_t seed,
                        uint64_t step, cudaStream_t st, int64_t *launches);
// CSVR: red3 = the {-, kinetic, n_mobile} triple launch_energy_reduce has just written; scales every velocity by lambda
void launch_csvr(int n_rows, float4 *vel, const double *red3, double kT, double c, double dof_removed, uint64_t seed, uint64_t step,
                 float *lambda, cudaStream_t st, int64_t *launches);
void launch_zero_velocities(int n_rows, float4 *vel, cudaStream_t st, int64_t *launches);
int com_partial_elems();
void launch_remove_com(int n_rows, float4 *vel, const uint8_t *flags, double *partial, cudaStream_t st, int64_t *launches);
void launch_scale_coords(int n, float4 *xyzq, float4 *xref, float4 *vel, float mu, float nu, cudaStream_t st, int64_t *launches);
// original-order <-> cell-order copies
void launch_gather_to_orig(int n, const float4 *sorted, const int *orig, float4 *out, cudaStream_t st, int64_t *launches);
void launch_scatter_from_orig(int n, const float4 *in_orig, const int *orig, float4 *sorted, int keep_w, cudaStream_t st,
                              int64_t *launches);
void launch_pack_xyz(int n, const float4 *sorted, const int *orig, float *out, cudaStream_t st, int64_t *launches);
void launch_l2_flush(float4 *buf, size_t n_float4, cudaStream_t st, int64_t *launches);
