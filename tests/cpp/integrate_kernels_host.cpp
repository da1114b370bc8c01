// Test infrastructure: the integrator kernels of integrate.cu (kick + drift with the displacement flag, the look-ahead
// margin, the non-finite guard and the host-published {tag, flag} word; the original-order gather / scatter) compiled
// unchanged for the host through tests/cpp/shim_mt/cuda_runtime.h.  tests/test_integrate_kernels_on_host.py compares
// them with the oracle's kick / drift.  The HALO instantiation (peer stores, system-scope flags) is not run this way:
// its protocol is model-checked in tests/test_halo_protocol_model.py and it is exercised on the GPUs.
#define MC_HOST_SHIM 1
#define MC_SHIM_SHARED_STATIC 1
#include "shim_mt/cuda_runtime.h"

static inline void __threadfence() {}
static inline void __threadfence_system() {}
static void halo_st_release_sys(uint32_t *, uint32_t) {}
static void halo_spin(const uint32_t *, uint32_t, int *) {}

#include "../../molchanica_b200/csrc/integrate.cu"

extern "C" {

void host_kick_drift(int n, float4 *xyzq, float4 *vel, const float4 *force, const float *ext_force, const int *orig, const uint8_t *flags,
                     const float4 *xref, float kick, float drift, float max_disp, float lookahead, int *rebuild_flag, int *host_flag,
                     int step_tag) {
    uint32_t done = 0;
    shim_launch((unsigned)((n + 255) / 256), 256, [&] {
        kick_drift_kernel<false>(n, xyzq, vel, force, ext_force, orig, flags, xref, kick, drift, max_disp, lookahead, rebuild_flag, HaloPush{},
                                 &done, host_flag, step_tag);
    });
}

void host_gather_scatter(int n, const float4 *sorted, const int *orig, float4 *out_orig, float4 *back, int keep_w) {
    shim_launch((unsigned)((n + 255) / 256), 256, [&] { gather_to_orig_kernel(n, sorted, orig, out_orig); });
    shim_launch((unsigned)((n + 255) / 256), 256, [&] { scatter_from_orig_kernel(n, out_orig, orig, back, keep_w); });
}
}
