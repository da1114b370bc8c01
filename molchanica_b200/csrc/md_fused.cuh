// md_fused.cuh -- host-side launcher of md_fused.cu (persistent multi-step kernel of small systems)
#pragma once
#include "bonded.cuh"
#include "common.cuh"

struct FusedArgs {
    int n;
    float4 *xyzq, *vel, *force;
    const float4 *xref;
    const uint16_t *type;
    const uint8_t *flags;
    const int *orig, *slot_of_orig;
    const uint32_t *nbr_start, *nbr_count, *nbr_list;  // global-slot rows
    const float2 *ljtab;
    NbParams p;
    int lj_on;
    const int32_t *p14_start, *p14_idx;  // nullptr: none
    float s14_lj, s14_q;
    BondedTerms bt;
    const float *ext_force;  // nullptr: none (constant over the call)
    float dt, max_disp;
    int n_steps, first_half;  // first_half: the first kick of this launch is a half kick (else a full one)
    int *rebuild_flag;        // bit 0: an atom moved more than max_disp since the list build, bit 1: non-finite coordinates
    int *out;                 // [0] drifts completed, [1] flag bits seen (pinned host memory)
};

int md_fused_max_atoms();
cudaError_t md_fused_prepare();
cudaError_t launch_md_fused(const FusedArgs &A, bool multi, int coul, bool pbc, int n_sms, cudaStream_t st, int64_t *launches);
