// Test infrastructure: the HOT kernels of pair_force.cu compiled unchanged for the host through the multi-threaded
// stand-in tests/cpp/shim_mt/cuda_runtime.h (threads of a block are OS threads, warp shuffles and __syncthreads are
// barriers) so that the pair-force path -- lane mapping, two-gathers-in-flight loop, interior / wrapped rows, warp
// reduction, multi-type table in "shared memory", the decomposed row order -- has a regression test that needs no GPU
// (tests/test_pair_kernel_on_host.py).  Not part of the product library; far too slow for anything but small systems.
#define MC_HOST_SHIM 1
#include "shim_mt/cuda_runtime.h"

static inline void __threadfence() {}
static void halo_spin(const uint32_t *, uint32_t, int *) {}



#include "../../molchanica_b200/csrc/pair_force.cu"

extern "C" {

// variant: 0 = <8, single type, no Coulomb, periodic, no energy>   (the C4 bench instantiation)
//          1 = <8, multi, plain Coulomb, vacuum, energy>            2 = <4, multi, erfc, periodic, energy>
//          3 = <16, multi, erfc, periodic, no energy, UNIFORM>      4 = <32, single, none, periodic, energy>
int host_pair_force(int variant, int n_rows, int row0, const float4 *xyzq, const uint16_t *type, const uint8_t *flags,
                    const uint32_t *nbr_start, const uint32_t *nbr_count, const uint32_t *nbr_list, const float2 *ljtab, int n_types,
                    const float *ext, int periodic, float rc_lj, float rc_q, float alpha, int lj_on, float4 *force, int n_interior,
                    int n_first) {
    NbParams p;
    memset(&p, 0, sizeof(p));
    for (int a = 0; a < 3; ++a) { p.ext[a] = periodic ? ext[a] : 1.f; p.inv_ext[a] = periodic ? 1.f / ext[a] : 1.f; }
    p.rc2_lj = rc_lj * rc_lj; p.rc2_q = rc_q * rc_q; p.alpha = alpha; p.periodic = periodic; p.n_types = n_types;
    p.sig2 = ljtab[0].x; p.eps24 = ljtab[0].y;
    const HaloWait hw{};
#define RUN(LANES, ...)                                                                                                   \
    shim_launch((unsigned)((n_rows + 128 / LANES - 1) / (128 / LANES)), 128, [&] {                                        \
        pair_force_kernel<LANES, __VA_ARGS__>(n_rows, row0, xyzq, type, flags, nbr_start, nbr_count, nbr_list, ljtab, p, lj_on, force, \
                                              n_interior, n_first, hw);                                                  \
    })
    switch (variant) {
        case 0: RUN(8, false, MC_COULOMB_NONE, true, false, false); break;
        case 1: RUN(8, true, MC_COULOMB_PLAIN, false, true, false); break;
        case 2: RUN(4, true, MC_COULOMB_ERFC, true, true, false); break;
        case 3: RUN(16, true, MC_COULOMB_ERFC, true, false, true); break;
        case 4: RUN(32, false, MC_COULOMB_NONE, true, true, false); break;
        default: return -1;
    }
#undef RUN
    return 0;
}

// pairs14_kernel (one thread per row; adds the scaled 1-4 terms to what the pair kernel wrote)
void host_pairs14(int n_rows, const float4 *xyzq, const uint16_t *type, const int *orig, const int *slot_of_orig, const int32_t *p14_start,
                  const int32_t *p14_idx, const float2 *ljtab, int n_types, const float *ext, int periodic, float scale_lj, float scale_q,
                  int lj_on, int coul_on, float4 *force) {
    NbParams p;
    memset(&p, 0, sizeof(p));
    for (int a = 0; a < 3; ++a) { p.ext[a] = periodic ? ext[a] : 1.f; p.inv_ext[a] = periodic ? 1.f / ext[a] : 1.f; }
    p.periodic = periodic; p.n_types = n_types;
    shim_launch((unsigned)((n_rows + 127) / 128), 128, [&] {
        pairs14_kernel(n_rows, 0, xyzq, type, orig, slot_of_orig, p14_start, p14_idx, ljtab, p, scale_lj, scale_q, lj_on, coul_on, force);
    });
}
}
