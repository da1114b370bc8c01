// vsite_terms.h -- massless virtual site of four-site water (OPC / TIP4P: the reference's md.water {o, h0, h1, m},
// properties/sol_shrinking_box.rs:605-613): M = O + a (H1 - O) + b (H2 - O), and the redistribution of the force
// that acts on M onto its three parents, which is the transpose of that linear map, so total force and torque
// are unchanged.  Shared by device (settle.cu) and host tests like bonded_terms.h.
#pragma once

#ifdef __CUDACC__
#define MC_VS_HD __host__ __device__ __forceinline__
#else
#define MC_VS_HD inline
#endif

// d1 = H1 - O, d2 = H2 - O (minimum image applied by the caller)
MC_VS_HD void mc_vsite_position(const float o[3], const float d1[3], const float d2[3], float a, float b, float m[3]) {
    for (int x = 0; x < 3; ++x) m[x] = o[x] + a * d1[x] + b * d2[x];
}

MC_VS_HD void mc_vsite_spread(const float fm[3], float a, float b, float fo[3], float f1[3], float f2[3]) {
    for (int x = 0; x < 3; ++x) {
        fo[x] = (1.f - a - b) * fm[x];
        f1[x] = a * fm[x];
        f2[x] = b * fm[x];
    }
}
