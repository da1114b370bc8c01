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
    int *out;                 // [0] drifts completed, [1] flag bits seen, [2] the host's list went stale, [3] in-kernel rebuilds (pinned host memory)
    // Brute-force mode (systems of at most md_fused_brute_max_atoms() atoms): the kernel keeps a PRIVATE Verlet list -- rows of
    // fixed stride, built by an all-pairs sweep (a thousand atoms: a microsecond) at the start of the launch and again, inside
    // the launch, whenever an atom has moved more than max_disp since -- so a call never comes back for a rebuild and the
    // host's list (sort, cells, tiles; ~0.3 ms at this size) is not needed for stepping at all.
    int brute;                // 0: rows of the host's list (nbr_*) and early exit on the displacement flag
    uint32_t *bl_list, *bl_count;  // [n x bl_stride], [n]
    uint32_t bl_stride;
    float4 *bl_xref;          // reference positions of the private list
    int *bl_flags;            // [0], [1]: displacement / non-finite bits of even / odd steps, [2]: host list stale, [3]: rebuilds (zeroed by the host)
    int bl_keep;              // 1: the private list of the previous launch is still good (no build at the start)
    int lanes;                // lanes per row in the force phase (set by the launcher)
    int need_forces;          // the forces of the starting positions are not in `force` yet: evaluate them first
    float rl2;                // squared list radius (cutoff + skin)
    const int32_t *excl_start, *excl_idx;  // nullptr: no exclusions (original ids)
    unsigned long long *dbg;  // MC_FUSED_TIMES=1: thread 0 stores %globaltimer at the phase boundaries of the launch (64 slots)
};

int md_fused_brute_max_atoms();

int md_fused_max_atoms();
cudaError_t md_fused_prepare();
cudaError_t launch_md_fused(const FusedArgs &A, bool multi, int coul, bool pbc, int n_sms, cudaStream_t st, int64_t *launches);
