// thermostat.cu -- Langevin thermostat on the device, see langevin_terms.h for the arithmetic.
// HBM-bound and tiny: 16 B read + 16 B written per atom, one launch per step when a thermostat is set.
#include "integrate.cuh"
#include "langevin_terms.h"
#include "csvr_terms.h"

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

// CSVR (csvr_terms.h): one thread turns the kinetic energy that energy_partial / energy_final have just reduced
// into the velocity scale factor of this step; csvr_scale_kernel applies it.  red3 = {-, sum 1/2 m v^2 in amu A^2/ps^2,
// mobile atoms}.  STATUS: arithmetic verified on the host (tests/test_csvr_cpu.py), kernels not yet run on hardware.
__global__ void csvr_lambda_kernel(const double *__restrict__ red3, double kT, double c, double dof_removed, uint64_t seed,
                                   uint64_t step, float *__restrict__ lambda_out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const double kinetic = red3[1] / 418.4;  // kcal/mol
    const double nf = 3.0 * red3[2] - dof_removed;
    *lambda_out = (float)mc_csvr_lambda(kinetic, kT, nf, c, seed, step);
}

__global__ void __launch_bounds__(256) csvr_scale_kernel(int n_rows, float4 *__restrict__ vel, const float *__restrict__ lambda) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const float l = *lambda;
    float4 v = vel[i];
    v.x *= l; v.y *= l; v.z *= l;
    vel[i] = v;
}

// velocities to zero, inverse masses kept: the quench between two steps of the energy minimiser (engine.cu)
__global__ void __launch_bounds__(256) zero_velocities_kernel(int n_rows, float4 *__restrict__ vel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    float4 v = vel[i];
    v.x = v.y = v.z = 0.f;
    vel[i] = v;
}

// MdConfig.zero_com_drift (reference properties/crystal.rs:310, water_sol.rs:144): remove the velocity of the centre of
// mass of the mobile atoms.  Two launches: per-block partial sums {sum m v, sum m} in fp64, then every block sums the
// (<= COM_BLOCKS) partials itself -- same order in every block, so all threads subtract the same vector -- and subtracts.
constexpr int COM_BLOCKS = 296;  // 2 x 148 SMs

__global__ void __launch_bounds__(256) com_partial_kernel(int n_rows, const float4 *__restrict__ vel, const uint8_t *__restrict__ flags,
                                                           double *__restrict__ partial) {
    double p[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += gridDim.x * blockDim.x) {
        const float4 v = vel[i];
        if (v.w > 0.f && !(flags[i] & MC_FLAG_STATIC)) {
            const double m = 1.0 / (double)v.w;
            p[0] += m * v.x; p[1] += m * v.y; p[2] += m * v.z; p[3] += m;
        }
    }
    __shared__ double sh[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) p[k] += __shfl_xor_sync(MC_FULL_MASK, p[k], d);
        if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = p[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
        partial[4 * blockIdx.x + threadIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) com_remove_kernel(int n_rows, float4 *__restrict__ vel, const uint8_t *__restrict__ flags,
                                                          const double *__restrict__ partial, int n_partial) {
    __shared__ float vcom[3];
    if (threadIdx.x == 0) {
        double t[4] = {0.0, 0.0, 0.0, 0.0};
        for (int b = 0; b < n_partial; ++b)
            for (int k = 0; k < 4; ++k) t[k] += partial[4 * b + k];
        for (int k = 0; k < 3; ++k) vcom[k] = t[3] > 0.0 ? (float)(t[k] / t[3]) : 0.f;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    float4 v = vel[i];
    if (v.w > 0.f && !(flags[i] & MC_FLAG_STATIC)) {
        v.x -= vcom[0]; v.y -= vcom[1]; v.z -= vcom[2];
        vel[i] = v;
    }
}

// barostat (engine.cu): positions and the displacement reference scaled about the origin by mu, velocities by nu
// (1 for Berendsen, 1 / mu for stochastic cell rescaling); charges / inverse masses in .w untouched
__global__ void __launch_bounds__(256) scale_coords_kernel(int n, float4 *__restrict__ xyzq, float4 *__restrict__ xref,
                                                            float4 *__restrict__ vel, float mu, float nu) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 x = xyzq[i], r = xref[i];
    x.x *= mu; x.y *= mu; x.z *= mu;
    r.x *= mu; r.y *= mu; r.z *= mu;
    xyzq[i] = x; xref[i] = r;
    if (nu != 1.f) {
        float4 v = vel[i];
        v.x *= nu; v.y *= nu; v.z *= nu;
        vel[i] = v;
    }
}

}  // namespace

#ifdef MC_HAVE_LAUNCH  // the stand-ins of tests/cpp/shim/ and shim_mt/ have no launcher; shim_fiber/ has
void launch_zero_velocities(int n_rows, float4 *vel, cudaStream_t st, int64_t *launches) {
    if (n_rows <= 0) return;
    MC_LAUNCH(zero_velocities_kernel, div_up(n_rows, 256), 256, 0, st, n_rows, vel);
    *launches += 1;
}

int com_partial_elems() { return 4 * COM_BLOCKS; }

void launch_remove_com(int n_rows, float4 *vel, const uint8_t *flags, double *partial, cudaStream_t st, int64_t *launches) {
    if (n_rows <= 0) return;
    const int nb = (int)min(div_up(n_rows, 256), (unsigned)COM_BLOCKS);
    MC_LAUNCH(com_partial_kernel, nb, 256, 0, st, n_rows, vel, flags, partial);
    MC_LAUNCH(com_remove_kernel, div_up(n_rows, 256), 256, 0, st, n_rows, vel, flags, partial, nb);
    *launches += 2;
}

void launch_scale_coords(int n, float4 *xyzq, float4 *xref, float4 *vel, float mu, float nu, cudaStream_t st, int64_t *launches) {
    if (n <= 0) return;
    MC_LAUNCH(scale_coords_kernel, div_up(n, 256), 256, 0, st, n, xyzq, xref, vel, mu, nu);
    *launches += 1;
}

void launch_csvr(int n_rows, float4 *vel, const double *red3, double kT, double c, double dof_removed, uint64_t seed, uint64_t step,
                 float *lambda, cudaStream_t st, int64_t *launches) {
    if (n_rows <= 0) return;
    MC_LAUNCH(csvr_lambda_kernel, 1, 32, 0, st, red3, kT, c, dof_removed, seed, step, lambda);
    MC_LAUNCH(csvr_scale_kernel, div_up(n_rows, 256), 256, 0, st, n_rows, vel, lambda);
    *launches += 2;
}

void launch_langevin_ou(int n_rows, float4 *vel, const int *orig, const uint8_t *flags, float c1, float c2, float kT, uint64_t seed,
                        uint64_t step, cudaStream_t st, int64_t *launches) {
    if (n_rows <= 0) return;
    MC_LAUNCH(langevin_ou_kernel, div_up(n_rows, 256), 256, 0, st, n_rows, vel, orig, flags, c1, c2, kT, seed, step);
    *launches += 1;
}

#endif
