// dock_filter.cu -- the clash pre-filter of process_poses (reference src/docking/legacy/mod.rs:522-573) on the device:
// one thread per pose poses the SAMPLED ligand carbons (every 4th ligand atom that is carbon, legacy/prep.rs:22) with
// the transform of pose_terms.h and tests them against the SAMPLED receptor carbons (every 6th near-site atom that is
// carbon, prep.rs:21,133): keep[p] = 0 when any pair is closer than 1.1 x the van der Waals radius.  The reference does
// this serially on the CPU before its rayon scoring loop; next to the 2.3 ms scoring kernel a host loop over 10k poses
// x ~800 x ~5 distances would dominate, here it is ~4e7 distance tests in one small launch (the sampled receptor, a
// few KB, stays in L1 / L2).  Host twin with identical arithmetic: mc_dock_filter_poses (dock_poses.cu).
// STATUS: kernel source verified on the host against that twin and the numpy restatement
// (tests/test_kernels_on_host.py); not yet run on hardware.
#include "dock.cuh"
#include "pose_terms.h"

namespace {

constexpr int DOCK_FILTER_MAX_LIG = 64;  // sampled ligand carbons per pose held in registers / local memory

__global__ void __launch_bounds__(128) dock_filter_kernel(int n_rs, const float4 *__restrict__ rec_sample, int n_ls,
                                                           const float4 *__restrict__ lig_sample, float3 anchor0, float limit,
                                                           int n_poses, const float *__restrict__ poses, uint8_t *__restrict__ keep) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_poses) return;
    const float *ps = poses + 7 * (size_t)p;
    const PoseQuat q = mc_pose_quat(ps);
    float lp[DOCK_FILTER_MAX_LIG][3];
    for (int a = 0; a < n_ls; ++a) {
        const float4 l = lig_sample[a];
        mc_pose_point(q, ps, l.x, l.y, l.z, anchor0.x, anchor0.y, anchor0.z, lp[a]);
    }
    bool clash = false;
    for (int r = 0; r < n_rs && !clash; ++r) {
        const float4 ra = rec_sample[r];
        for (int a = 0; a < n_ls; ++a) {
            const float ex = ra.x - lp[a][0], ey = ra.y - lp[a][1], ez = ra.z - lp[a][2];
            if (sqrtf(ex * ex + ey * ey + ez * ez) < limit) { clash = true; break; }
        }
    }
    keep[p] = clash ? 0 : 1;
}

}  // namespace

int dock_filter_max_lig() { return DOCK_FILTER_MAX_LIG; }

#ifdef MC_HAVE_LAUNCH  // the stand-ins of tests/cpp/shim/ and shim_mt/ have no launcher; shim_fiber/ has
void launch_dock_filter(int n_rs, const float4 *rec_sample, int n_ls, const float4 *lig_sample, float3 anchor0, float limit, int n_poses,
                        const float *poses, uint8_t *keep, cudaStream_t st, int64_t *launches) {
    if (n_poses <= 0) return;
    MC_LAUNCH(dock_filter_kernel, div_up((size_t)n_poses, 128), 128, 0, st, n_rs, rec_sample, n_ls, lig_sample, anchor0, limit, n_poses, poses, keep);
    *launches += 1;
}
#endif
