// bonded.cu -- bonded forces on the device: harmonic bonds and angles, Amber periodic
// dihedrals, added to the nonbonded force of the same evaluation so that the whole step stays on the GPU
// (the reference's MdState::step evaluates bonded + nonbonded inside one call, README.md:234-241).
//
// One thread per term; the arithmetic is bonded_terms.h (verified on the host by finite differences,
// tests/cpp/bonded_math_check.cpp).  Terms are stored with the caller's atom ids and resolved to the current
// cell-order slots through slot_of_orig at every launch (the order changes at each list build).  Forces are
// accumulated with fp32 atomics into the float4 force array the pair kernel has just written (a few terms
// per atom; the sum order is not fixed, which is within the 1e-5 parity bar), energies into one fp64 word.
// HBM-bound and tiny next to the pair kernel: 3 x 16 B gathered + 3 x 12 B of atomics per bond.
// STATUS: written after round 1's GPU budget was spent -- compiled for sm_100a, arithmetic verified on the host,
// kernel plumbing not yet run on hardware (tests/test_gpu_bonded.py is marked accordingly).
#include "bonded_device.cuh"

namespace {

__global__ void __launch_bounds__(128) bonded_kernel(BondedTerms t, const int *__restrict__ slot_of_orig,
                                                      const float4 *__restrict__ xyzq, const NbParams p,
                                                      float4 *__restrict__ force, double *__restrict__ energy3,
                                                      int want_energy) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float e, w;  // energy and sum_a (r_a - r_ref) . f_a of this thread's term (its share of the virial)
    int kind;
    bonded_term_apply(tid, t, slot_of_orig, xyzq, p, force, e, w, kind);
    if (want_energy) {
        // per-kind energy sums: warp-reduce, one fp64 atomic per warp and kind present in it
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float v = kind == k ? e : 0.f;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(MC_FULL_MASK, v, d);
            if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(energy3 + k, (double)v);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(MC_FULL_MASK, w, d);
        if ((threadIdx.x & 31) == 0 && w != 0.f) atomicAdd(energy3 + 3, (double)w);  // [3]: virial of the bonded terms
    }
}

}  // namespace

#ifdef MC_HAVE_LAUNCH  // the serial stand-in of tests/cpp/shim/ has no launcher
void launch_bonded(const BondedTerms &t, const int *slot_of_orig, const float4 *xyzq, const NbParams &p, float4 *force,
                   double *energy3, bool want_energy, cudaStream_t st, int64_t *launches) {
    const int n = t.n_bonds + t.n_angles + t.n_dihedrals;
    if (n <= 0) return;
    if (want_energy) cudaMemsetAsync(energy3, 0, 4 * sizeof(double), st);
    MC_LAUNCH(bonded_kernel, div_up((size_t)n, 128), 128, 0, st, t, slot_of_orig, xyzq, p, force, energy3, want_energy ? 1 : 0);
    *launches += 1;
}
#endif
