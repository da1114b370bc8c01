// Test infrastructure: the device prefix scan and the stable LSD radix sort of sort_scan.cu -- kernels AND their host
// drivers (exclusive_scan_u32, radix_sort_pairs: pass loop, scratch layout, ping-pong buffers) -- compiled unchanged for
// the host through tests/cpp/shim_mt/cuda_runtime.h (MC_LAUNCH of common.cuh becomes a run of the blocks on OS threads;
// __match_any_sync / __ballot_sync / __shfl_up_sync are emulated).  tests/test_sort_scan_on_host.py compares them with
// numpy (stable argsort, cumsum).  Not part of the product library.
#define MC_HOST_SHIM 1
#define MC_SHIM_SHARED_STATIC 1
#include "shim_mt/cuda_runtime.h"

#include "../../molchanica_b200/csrc/sort_scan.cu"

extern "C" {
// out has n + 1 elements (the grand total goes to out[n]); in == out is allowed, as radix_sort_pairs uses it
void host_exclusive_scan(const uint32_t *in, uint32_t *out, size_t n, int align8, int64_t *launches) {
    std::vector<uint32_t> scratch(scan_scratch_elems(n) + 8);
    exclusive_scan_u32(in, out, n, align8, scratch.data(), nullptr, launches);
}

// keys0 / vals0 hold the input; returns which of the two buffer pairs holds the sorted result
int host_radix_sort_pairs(uint32_t *keys0, uint32_t *vals0, uint32_t *keys1, uint32_t *vals1, size_t n, int bits, int64_t *launches) {
    std::vector<uint32_t> scratch(radix_scratch_elems(n) + 8);
    uint32_t *keys[2] = {keys0, keys1}, *vals[2] = {vals0, vals1};
    return radix_sort_pairs(keys, vals, n, bits, scratch.data(), nullptr, launches);
}
}
