// dock.cuh -- host-side launcher of dock.cu
#pragma once
#include "common.cuh"

// meta words: LJ type in bits 0-15, hydrophobic flag in bit 16
void launch_dock_score(int n_rec, const float4 *rec, const uint32_t *rec_meta, int n_lig, const float4 *lig,
                       const uint32_t *lig_meta, float3 lig_anchor, int n_rec_types, int n_lig_types,
                       const float2 *ljtab /* (sigma^2, 4 eps) */, int n_poses, const float *poses, float *out,
                       cudaStream_t st, int64_t *launches);
size_t dock_smem_bytes(int n_lig, int n_rec_types, int n_lig_types);
cudaError_t dock_prepare();
