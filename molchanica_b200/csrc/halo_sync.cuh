// halo_sync.cuh -- device side of the peer-memory ghost exchange that is fused into the
// step kernels: kick_drift stores the new positions of its boundary layers straight into the
// neighbour ranks' ghost blocks through NVLink-mapped peer pointers (cudaIpc), the last block to
// finish raises a "ready" flag in the peers' memory, and the pair kernel of the boundary rows spins
// on its own copy of that flag before it reads a ghost.  No pack / unpack, no NCCL group, no extra
// launch on the step path.
//
// Flags are monotonically increasing step epochs.  Per rank, in its own (exported) memory:
//   ready_from_{prev,next}  written by that neighbour when its push of epoch e has landed
//   ack_from_{prev,next}    written by that neighbour when every kernel of it that read the ghosts
//                           of epoch <= e has finished (so the next push may overwrite them)
// Dependencies only point backwards in time (push e+1 waits for ack e, pair e waits for ready e),
// so there is no cycle; every spin is bounded by a timeout that raises bit 2 of the error word
// instead of hanging the device when a peer died.
#pragma once
#include <stdint.h>

#define MC_HALO_READY_FROM_PREV 0
#define MC_HALO_READY_FROM_NEXT 8
#define MC_HALO_ACK_FROM_PREV 16
#define MC_HALO_ACK_FROM_NEXT 24
#define MC_HALO_CNT_KICK 32
#define MC_HALO_FLAG_WORDS 64
#define MC_HALO_ERR_TIMEOUT 4

// Push side, passed by value to kick_drift_kernel<true>.
struct HaloPush {
    float4 *to_prev, *to_next;   // peer ghost blocks (mapped peer pointers); nullptr = no push this step
    int n_first, last_begin;     // rows [0, n_first) go to prev, rows [last_begin, n_rows) to next
    const uint32_t *ack_prev, *ack_next;  // local flags the peers write
    uint32_t *sig_ack_prev, *sig_ack_next;      // peers' ack_from_{next,prev} words
    uint32_t *sig_ready_prev, *sig_ready_next;  // peers' ready_from_{next,prev} words
    uint32_t *done_counter;      // local
    uint32_t epoch;
    int *err;
    int *max_disp2;  // float bits of the largest squared displacement since the last build (adaptive rebuild interval)
};

// Wait side, read by the pair kernel of the boundary rows.
struct HaloWait {
    const uint32_t *ready_prev, *ready_next;  // nullptr = no wait
    uint32_t want;
    int *err;
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t halo_ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void halo_st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long halo_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// One thread: wait until *flag >= want (wrap-safe), at most ~2 s.
static __device__ __noinline__ void halo_spin(const uint32_t *flag, uint32_t want, int *err) {
    if ((int)(halo_ld_acquire_sys(flag) - want) >= 0) return;
    const unsigned long long t0 = halo_timer_ns();
    while ((int)(halo_ld_acquire_sys(flag) - want) < 0) {
        if ((*(volatile int *)err & MC_HALO_ERR_TIMEOUT) != 0) return;  // somebody already gave up
        if (halo_timer_ns() - t0 > 2000000000ull) {
            atomicOr(err, MC_HALO_ERR_TIMEOUT);
            return;
        }
        __nanosleep(40);
    }
}
#elif defined(MC_HOST_LAUNCH)
// host build of tests/cpp/host_lib/: a single process has no peers, the fused-halo kernels are compiled but never launched
#include <atomic>
static inline uint32_t halo_ld_acquire_sys(const uint32_t *p) { return std::atomic_ref<const uint32_t>(*p).load(std::memory_order_acquire); }
static inline void halo_st_release_sys(uint32_t *p, uint32_t v) { std::atomic_ref<uint32_t>(*p).store(v, std::memory_order_release); }
static inline void halo_spin(const uint32_t *flag, uint32_t want, int *err) {
    for (long k = 0; (int)(halo_ld_acquire_sys(flag) - want) < 0; ++k) {
        if (k > 100000000L) { atomicOr(err, MC_HALO_ERR_TIMEOUT); return; }
        __nanosleep(40);
    }
}
#endif
