// Test infrastructure: the pose-energy scan kernel of dock.cu (one block per pose: f64 pose transform into shared memory,
// the receptor streamed by the block's threads, block reduction of the five terms) compiled unchanged for the host
// through tests/cpp/shim_mt/cuda_runtime.h.  tests/test_dock_kernel_on_host.py holds it to the GPU parity bar against
// the fp64 oracle.  Not part of the product library.
#define MC_HOST_SHIM 1
#define MC_SHIM_SHARED_STATIC 1
#include "shim_mt/cuda_runtime.h"

#include "../../molchanica_b200/csrc/dock.cu"

extern "C" void host_dock_score(int n_rec, const float4 *rec, const uint32_t *rec_meta, int n_lig, const float4 *lig, const uint32_t *lig_meta,
                                const float *anchor, int n_rec_types, int n_lig_types, const float2 *ljtab, int n_poses, const float *poses,
                                float *out) {
    shim_launch((unsigned)n_poses, DOCK_THREADS, [&] {
        dock_score_kernel(n_rec, rec, rec_meta, n_lig, lig, lig_meta, make_float3(anchor[0], anchor[1], anchor[2]), n_rec_types, n_lig_types,
                          ljtab, poses, 7, 0, nullptr, nullptr, out);
    });
}
