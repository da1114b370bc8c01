// langevin_terms.h -- arithmetic of the Langevin thermostat (the reference offers CSVR and Langevin
// thermostats, README.md:238, ui/panels/md.rs:296-306,433-556): the Ornstein-Uhlenbeck velocity update
//     v <- c1 v + c2 sqrt(kT / m) xi,   c1 = exp(-gamma dt),  c2 = sqrt(1 - c1^2),  xi ~ N(0, 1)
// with counter-based noise (Philox4x32-10, Salmon et al., SC'11): the three normals of atom `id` at step `step`
// are a pure function of (seed, id, step), so the result does not depend on how atoms are ordered or spread over
// threads, and the CPU oracle can draw the very same numbers.  Shared by device and host like bonded_terms.h;
// tests/test_langevin_cpu.py checks the generator against the published known-answer vectors.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define MC_LGV_HD __host__ __device__ __forceinline__
#else
#define MC_LGV_HD inline
#endif

MC_LGV_HD void mc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// (0, 1] uniform from 32 random bits
MC_LGV_HD float mc_u01(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }

// Three standard normals of (seed, atom id, step): Box-Muller on the four Philox words.
MC_LGV_HD void mc_langevin_normals(uint64_t seed, uint32_t atom_id, uint64_t step, float xi[3]) {
    const uint32_t ctr[4] = {atom_id, (uint32_t)step, (uint32_t)(step >> 32), 0u};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t r[4];
    mc_philox4x32_10(ctr, key, r);
    const float ra = sqrtf(-2.0f * logf(mc_u01(r[0]))), ta = 6.2831853071795865f * mc_u01(r[1]);
    const float rb = sqrtf(-2.0f * logf(mc_u01(r[2]))), tb = 6.2831853071795865f * mc_u01(r[3]);
    xi[0] = ra * cosf(ta);
    xi[1] = ra * sinf(ta);
    xi[2] = rb * cosf(tb);
}

// One Ornstein-Uhlenbeck update of a velocity (A/ps); inv_mass in 1/amu, kT in kcal/mol (418.4 A^2/ps^2 per kcal/mol/amu).
MC_LGV_HD void mc_langevin_ou(float v[3], float inv_mass, float c1, float c2, float kT, const float xi[3]) {
    const float s = c2 * sqrtf(kT * inv_mass * 418.4f);
    v[0] = c1 * v[0] + s * xi[0];
    v[1] = c1 * v[1] + s * xi[1];
    v[2] = c1 * v[2] + s * xi[2];
}
