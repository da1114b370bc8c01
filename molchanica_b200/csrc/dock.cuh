// dock.cuh -- host-side launcher of dock.cu
#pragma once
#include "common.cuh"

// meta words: LJ type in bits 0-15, hydrophobic flag in bit 16
void launch_dock_score(int n_rec, const float4 *rec, const uint32_t *rec_meta, int n_lig, const float4 *lig,
                       const uint32_t *lig_meta, float3 lig_anchor, int n_rec_types, int n_lig_types,
                       const float2 *ljtab /* (sigma^2, 4 eps) */, int n_poses, const float *poses, int pose_stride /* 7 + n_flex */,
                       int n_flex, const int2 *flex_axis, const uint8_t *flex_mask /* [n_flex][n_lig] */, float *out,
                       cudaStream_t st, int64_t *launches);
size_t dock_smem_bytes(int n_lig, int n_rec_types, int n_lig_types);
cudaError_t dock_prepare();

// dock_filter.cu: keep[p] = 0 when a sampled ligand carbon of pose p comes within `limit` of a sampled receptor carbon
int dock_filter_max_lig();
void launch_dock_filter(int n_rs, const float4 *rec_sample, int n_ls, const float4 *lig_sample, float3 anchor0, float limit, int n_poses,
                        const float *poses, uint8_t *keep, cudaStream_t st, int64_t *launches);
