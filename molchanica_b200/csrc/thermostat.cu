// thermostat.cu -- Langevin thermostat on the device, see langevin_terms.h for the arithmetic.
// HBM-bound and tiny: 16 B read + 16 B written per atom, one launch per step when a thermostat is set.
#include "integrate.cuh"
#include "langevin_terms.h"

namespace {

// Langevin thermostat, O step (langevin_terms.h): applied after the drift (and the constraints) of a step, before
// its force evaluation -- the splitting B A O B.  Noise is keyed by the atom's ORIGINAL id and the step counter.
// STATUS: arithmetic verified on the host (tests/test_langevin_cpu.py), kernel not yet run on hardware.
__global__ void __launch_bounds__(256) langevin_ou_kernel(int n_rows, float4 *__restrict__ vel, const int *__restrict__ orig,
                                                           const uint8_t *__restrict__ flags, float c1, float c2, float kT,
                                                           uint64_t seed, uint64_t step) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows || (flags[i] & MC_FLAG_STATIC)) return;
    float4 v = vel[i];
    if (v.w <= 0.f) return;
    float xi[3], vv[3] = {v.x, v.y, v.z};
    mc_langevin_normals(seed, (uint32_t)orig[i], step, xi);
    mc_langevin_ou(vv, v.w, c1, c2, kT, xi);
    v.x = vv[0]; v.y = vv[1]; v.z = vv[2];
    vel[i] = v;
}

}  // namespace

#ifndef MC_HOST_SHIM
void launch_langevin_ou(int n_rows, float4 *vel, const int *orig, const uint8_t *flags, float c1, float c2, float kT, uint64_t seed,
                        uint64_t step, cudaStream_t st, int64_t *launches) {
    if (n_rows <= 0) return;
    langevin_ou_kernel<<<div_up(n_rows, 256), 256, 0, st>>>(n_rows, vel, orig, flags, c1, c2, kT, seed, step);
    *launches += 1;
}

#endif
