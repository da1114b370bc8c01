// settle.cuh -- host-side launcher of settle.cu
#pragma once
#include "common.cuh"

// waters: (O, H1, H2, -) original ids.  Constrains the positions kick_drift advanced by `dt` and corrects the velocities.
void launch_settle(int n_w, const int4 *waters, const int *slot_of_orig, float4 *xyzq, float4 *vel, float m_o, float m_h,
                   float d_oh, float d_hh, const NbParams &p, float dt, double *virial, cudaStream_t st, int64_t *launches);

// Virtual sites M = O + a (H1 - O) + b (H2 - O); sites: (M, O, H1, H2) original ids.
void launch_vsite_construct(int n_v, const int4 *sites, const int *slot_of_orig, float4 *xyzq, float a, float b, const NbParams &p,
                            cudaStream_t st, int64_t *launches);
void launch_vsite_spread(int n_v, const int4 *sites, const int *slot_of_orig, float4 *force, float a, float b, cudaStream_t st,
                         int64_t *launches);

// SHAKE for bonds to hydrogen: clusters (heavy, h1, h2, h3) original ids (-1 = unused), dist 3 lengths per cluster;
// *not_converged (device) counts clusters that needed more than 64 sweeps.
void launch_shake_h(int n_c, const int4 *clusters, const float *dist, const int *slot_of_orig, float4 *xyzq, float4 *vel,
                    const NbParams &p, float dt, float tol, int *not_converged, double *virial, cudaStream_t st, int64_t *launches);

// Velocity stage of RATTLE after a closing half kick: removes the velocity components along the constrained bonds of the
// rigid waters and of the hydrogen clusters (either count may be 0).
void launch_rattle_velocities(int n_w, const int4 *waters, int n_c, const int4 *clusters, const int *slot_of_orig, const float4 *xyzq,
                              float4 *vel, const NbParams &p, cudaStream_t st, int64_t *launches);
