// Test infrastructure: host stand-ins for the asynchronous-copy machinery tile_build.cu uses on sm_100a -- an mbarrier
// (arrival count + transaction bytes + phase parity), the 1-D bulk copy that completes on it, and a named barrier of a
// subset of the block's threads.  Semantics as in the PTX ISA: a phase completes when the pending arrival count AND the
// outstanding transaction bytes are both zero; completion flips the phase parity and re-arms the count; try_wait.parity(P)
// succeeds once the phase of parity P has completed.  One mutex + condition variable for everything: speed is no goal.
#pragma once
#include <condition_variable>
#include <map>
#include <mutex>

struct ShimMbar { uint32_t init = 0, pending = 0, phase = 0; int64_t tx = 0; };
static std::mutex shim_mbar_mu;
static std::condition_variable shim_mbar_cv;
static std::map<const void *, ShimMbar> shim_mbars;

static inline void shim_mbar_settle(ShimMbar &b) {  // caller holds the mutex
    if (b.pending == 0 && b.tx == 0) {
        b.phase ^= 1u;
        b.pending = b.init;
        shim_mbar_cv.notify_all();
    }
}
static inline void shim_mbar_init(uint64_t *bar, uint32_t count) {
    std::lock_guard<std::mutex> lk(shim_mbar_mu);
    ShimMbar &b = shim_mbars[bar];
    b = ShimMbar{};
    b.init = b.pending = count;
}
// mbarrier.arrive (tx_bytes == 0) / mbarrier.arrive.expect_tx
static inline void shim_mbar_arrive(uint64_t *bar, uint32_t tx_bytes) {
    std::lock_guard<std::mutex> lk(shim_mbar_mu);
    ShimMbar &b = shim_mbars.at(bar);
    b.tx += tx_bytes;
    b.pending -= 1;
    shim_mbar_settle(b);
}
// cp.async.bulk ... mbarrier::complete_tx::bytes.  The hardware requires 16-byte alignment and sizes: hold the caller to it.
static inline void shim_bulk_copy(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | bytes) & 15u) abort();
    memcpy(dst, src, bytes);
    std::lock_guard<std::mutex> lk(shim_mbar_mu);
    ShimMbar &b = shim_mbars.at(bar);
    b.tx -= bytes;
    shim_mbar_settle(b);
}
static inline void shim_mbar_wait(uint64_t *bar, uint32_t parity) {
    std::unique_lock<std::mutex> lk(shim_mbar_mu);
    ShimMbar &b = shim_mbars.at(bar);
    shim_mbar_cv.wait(lk, [&] { return b.phase != (parity & 1u); });
}

// bar.sync id, n_threads
struct ShimNamedBar { unsigned count = 0, gen = 0; };
static ShimNamedBar shim_named_bars[16];
static inline void shim_named_barrier(int id, unsigned n_threads) {
    std::unique_lock<std::mutex> lk(shim_mbar_mu);
    ShimNamedBar &b = shim_named_bars[id];
    const unsigned g = b.gen;
    if (++b.count == n_threads) {
        b.count = 0;
        ++b.gen;
        shim_mbar_cv.notify_all();
    } else {
        shim_mbar_cv.wait(lk, [&] { return b.gen != g; });
    }
}
