/*
 * ref_cuda_host.cpp -- builds the reference's own src/cuda/cuda.cu (+ util.cu, which it
 * includes) for the host and exports its __device__ helpers through a C ABI so the oracle can
 * be pinned against them.  The reference sources are #included from where they lie under
 * /root/reference (path given by -DREF_CUDA_CU=...); nothing is copied into this repository.
 * The extern "C" kernels of cuda.cu (coulomb_force_kernel, lj_V_kernel, lj_force_kernel)
 * become ordinary host functions under the shim and are exported as they are.
 */
#include "cuda_host_shim.h"
#include REF_CUDA_CU

extern "C" {

/* util.cu:119-139 */
void ref_lj_force(const float *tgt, const float *src, float sigma, float eps, float *out4) {
    ForceEnergy fe = lj_force(make_float3(tgt[0], tgt[1], tgt[2]), make_float3(src[0], src[1], src[2]), sigma, eps);
    out4[0] = fe.force.x; out4[1] = fe.force.y; out4[2] = fe.force.z; out4[3] = fe.energy;
}

/* util.cu:74-90 */
float ref_lj_V(const float *p0, const float *p1, float sigma, float eps) {
    return lj_V(make_float3(p0[0], p0[1], p0[2]), make_float3(p1[0], p1[1], p1[2]), sigma, eps);
}

/* util.cu:54-63 */
void ref_coulomb_force(const float *src, const float *tgt, float q_src, float q_tgt, float *out3) {
    float3 f = coulomb_force(make_float3(src[0], src[1], src[2]), make_float3(tgt[0], tgt[1], tgt[2]), q_src, q_tgt);
    out3[0] = f.x; out3[1] = f.y; out3[2] = f.z;
}

/* util.cu:65-71 */
void ref_min_image(const float *ext, const float *dv, float *out3) {
    float3 r = min_image(make_float3(ext[0], ext[1], ext[2]), make_float3(dv[0], dv[1], dv[2]));
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

float ref_softening_sq(void) { return SOFTENING_FACTOR_SQ; }
float ref_inv_sqrt_pi(void) { return INV_SQRT_PI; }

}
