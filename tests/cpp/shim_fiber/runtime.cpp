// Test infrastructure: the execution engine behind tests/cpp/shim_fiber/cuda_runtime.h.  See the header for what it is for.
//
// A block runs on ONE OS thread: its threads are fibers with their own stacks, switched by a dozen instructions of
// assembly (callee-saved registers + stack pointer).  Scheduling is cooperative and two-level: a thread that waits for
// its warp (shuffle, vote, __syncwarp) passes control to the next live lane of the SAME warp; a thread that waits for
// anything else (__syncthreads, named barrier, mbarrier, a spin on memory) passes control to the NEXT warp.  On return a
// warp resumes one lane past the lane that left it, so no lane starves.  A thread that exits counts as arrived at every
// barrier it no longer reaches.  The blocks of a launch are handed out to a small pool of OS threads.
#include "cuda_runtime.h"
#include "cuda.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

thread_local ShimThreadState shim_ts = {{0, 0, 0}, {0, 0, 0}, {1, 1, 1}, {1, 1, 1}, nullptr};

// ---- context switch (x86-64 System V) ---------------------------------------------------------------------------------------------
extern "C" void shim_ctx_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl shim_ctx_switch
.type shim_ctx_switch,@function
shim_ctx_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size shim_ctx_switch,.-shim_ctx_switch
)");

namespace {

constexpr unsigned MAX_THREADS = 1024;
constexpr size_t STACK_BYTES = 256 * 1024;
constexpr size_t DYN_SMEM_BYTES = 256 * 1024;

struct Fiber {
    void *sp = nullptr;
    bool done = true;
};

struct Mbar { uint32_t init = 0, pending = 0, phase = 0; int64_t tx = 0; };

struct BlockRun {
    unsigned nthreads = 0, nwarps = 0, cur = 0, live_block = 0;
    Fiber fibers[MAX_THREADS];
    char *stacks = nullptr;  // MAX_THREADS x STACK_BYTES, committed lazily by the kernel
    unsigned live_warp[32], resume_lane[32];
    unsigned sync_cnt = 0, sync_gen = 0;
    unsigned wb_cnt[32], wb_gen[32];
    unsigned mb_cnt[32], mb_gen[32];
    unsigned nb_cnt[16], nb_gen[16];
    uint64_t xchg[MAX_THREADS];
    void *main_sp = nullptr;
    const std::function<void()> *body = nullptr;
    uint64_t idle_switches = 0;  // consecutive switches without progress: deadlock watchdog
    std::map<const void *, Mbar> mbars;
};

thread_local BlockRun *run = nullptr;

void progress() { run->idle_switches = 0; }

void switch_to(unsigned next) {
    BlockRun &r = *run;
    const unsigned prev = r.cur;
    if (next == prev) return;
    if (++r.idle_switches > 400000000ull) {
        fprintf(stderr, "shim_fiber: no progress after 4e8 context switches in block %u (deadlock in the kernel under test?)\n", shim_ts.bidx.x);
        abort();
    }
    r.cur = next;
    shim_ts.tidx.x = next;
    shim_ctx_switch(&r.fibers[prev].sp, r.fibers[next].sp);
}

// next live lane of the same warp, or `t` itself
unsigned next_in_warp(unsigned t) {
    BlockRun &r = *run;
    const unsigned w = t >> 5, lane = t & 31u;
    for (unsigned l = 1; l < 32; ++l) {
        const unsigned c = (w << 5) | ((lane + l) & 31u);
        if (c < r.nthreads && !r.fibers[c].done) return c;
    }
    return t;
}

// a live thread of another warp (resuming that warp where it was left), or MAX_THREADS if there is none
unsigned next_warp_thread(unsigned t) {
    BlockRun &r = *run;
    const unsigned w = t >> 5;
    for (unsigned dw = 1; dw < r.nwarps; ++dw) {
        const unsigned w2 = (w + dw) % r.nwarps;
        if (r.live_warp[w2] == 0) continue;
        for (unsigned l = 0; l < 32; ++l) {
            const unsigned c = (w2 << 5) | ((r.resume_lane[w2] + l) & 31u);
            if (c < r.nthreads && !r.fibers[c].done) return c;
        }
    }
    return MAX_THREADS;
}

void yield_lane() { switch_to(next_in_warp(run->cur)); }

void yield_block() {
    BlockRun &r = *run;
    const unsigned t = r.cur, nxt = next_warp_thread(t);
    if (nxt == MAX_THREADS) { yield_lane(); return; }
    r.resume_lane[t >> 5] = next_in_warp(t) & 31u;
    switch_to(nxt);
}

void release_if_complete_warp(unsigned w) {
    BlockRun &r = *run;
    if (r.wb_cnt[w] > 0 && r.wb_cnt[w] >= r.live_warp[w]) { r.wb_cnt[w] = 0; ++r.wb_gen[w]; progress(); }
}
void release_if_complete_block() {
    BlockRun &r = *run;
    if (r.sync_cnt > 0 && r.sync_cnt >= r.live_block) { r.sync_cnt = 0; ++r.sync_gen; progress(); }
}

void fiber_exit() {
    BlockRun &r = *run;
    const unsigned t = r.cur, w = t >> 5;
    r.fibers[t].done = true;
    --r.live_block;
    --r.live_warp[w];
    progress();
    release_if_complete_warp(w);
    release_if_complete_block();
    unsigned nxt = next_in_warp(t);
    if (nxt == t) nxt = next_warp_thread(t);
    if (nxt == t || nxt == MAX_THREADS) {  // last thread of the block
        shim_ctx_switch(&r.fibers[t].sp, r.main_sp);
    } else {
        r.cur = nxt;
        shim_ts.tidx.x = nxt;
        shim_ctx_switch(&r.fibers[t].sp, r.fibers[nxt].sp);
    }
    abort();  // a finished fiber is never resumed
}

void fiber_entry() {
    (*run->body)();
    fiber_exit();
}

BlockRun *thread_run() {
    static thread_local BlockRun *mine = nullptr;
    if (!mine) {
        mine = new BlockRun();
        mine->stacks = static_cast<char *>(mmap(nullptr, MAX_THREADS * STACK_BYTES, PROT_READ | PROT_WRITE,
                                                MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
        if (mine->stacks == MAP_FAILED) { perror("shim_fiber: mmap"); abort(); }
    }
    if (!shim_ts.dyn_smem) shim_ts.dyn_smem = static_cast<unsigned char *>(aligned_alloc(1024, DYN_SMEM_BYTES));
    return mine;
}

void run_block(unsigned b, unsigned threads, const std::function<void()> &body) {
    BlockRun &r = *thread_run();
    run = &r;
    r.nthreads = threads;
    r.nwarps = (threads + 31) / 32;
    r.live_block = threads;
    r.sync_cnt = 0;
    r.body = &body;
    r.idle_switches = 0;
    r.mbars.clear();
    for (unsigned w = 0; w < 32; ++w) {
        r.live_warp[w] = w < r.nwarps ? ((w + 1) * 32 <= threads ? 32 : threads - w * 32) : 0;
        r.resume_lane[w] = 0;
        r.wb_cnt[w] = r.mb_cnt[w] = 0;
    }
    for (unsigned i = 0; i < 16; ++i) r.nb_cnt[i] = 0;
    for (unsigned t = 0; t < threads; ++t) {
        // initial frame: six callee-saved registers, the entry point as return address, a null return address above it
        uintptr_t top = reinterpret_cast<uintptr_t>(r.stacks + (size_t)(t + 1) * STACK_BYTES) & ~(uintptr_t)15;
        void **sp = reinterpret_cast<void **>(top);
        *--sp = nullptr;                                   // top - 8: fake return address of fiber_entry
        *--sp = reinterpret_cast<void *>(&fiber_entry);    // top - 16: popped by `ret`
        for (int k = 0; k < 6; ++k) *--sp = nullptr;       // rbp rbx r12 r13 r14 r15
        r.fibers[t].sp = sp;
        r.fibers[t].done = false;
    }
    shim_ts.bidx = {b, 0, 0};
    shim_ts.tidx = {0, 0, 0};
    r.cur = 0;
    shim_ctx_switch(&r.main_sp, r.fibers[0].sp);
    if (r.live_block != 0) { fprintf(stderr, "shim_fiber: block %u ended with %u live threads\n", b, r.live_block); abort(); }
    run = nullptr;
}

// ---- pool of OS threads that share the blocks of a launch ------------------------------------------------------------------------------------
struct Pool {
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::vector<std::thread> workers;
    uint64_t job_id = 0;
    unsigned blocks = 0, threads = 0, active = 0;
    std::atomic<unsigned> next{0};
    const std::function<void()> *body = nullptr;
    bool stop = false;

    void work(unsigned grid) {
        shim_ts.bdim = {threads, 1, 1};
        shim_ts.gdim = {grid, 1, 1};
        for (;;) {
            const unsigned b = next.fetch_add(1);
            if (b >= grid) break;
            run_block(b, threads, *body);
        }
    }
    void worker_main() {
        uint64_t seen = 0;
        for (;;) {
            unsigned grid;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_job.wait(lk, [&] { return stop || job_id != seen; });
                if (stop) return;
                seen = job_id;
                grid = blocks;
            }
            work(grid);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--active == 0) cv_done.notify_all();
            }
        }
    }
    void launch(unsigned grid, unsigned nthreads, const std::function<void()> &fn) {
        static const unsigned n_workers = [] {
            const char *e = getenv("MC_SHIM_THREADS");
            unsigned n = e ? (unsigned)atoi(e) : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
            return n < 1 ? 1u : n;
        }();
        const unsigned helpers = grid >= 4 ? std::min(n_workers - 1, grid - 1) : 0;  // the caller works too
        if (helpers == 0) {
            threads = nthreads; body = &fn; next = 0;
            work(grid);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            while (workers.size() < n_workers - 1) workers.emplace_back([this] { worker_main(); });
            blocks = grid; threads = nthreads; body = &fn; next = 0;
            active = (unsigned)workers.size();
            ++job_id;
        }
        cv_job.notify_all();
        work(grid);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return active == 0; });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_job.notify_all();
        for (auto &t : workers) t.join();
    }
};

Pool &pool() {
    static Pool *p = new Pool();  // leaked on purpose: no join at process exit from a ctypes-loaded library
    return *p;
}

int g_last_error = 0;

}  // namespace

void shim_launch(unsigned blocks, unsigned threads, const std::function<void()> &body) {
    if (blocks == 0) return;
    if (threads == 0 || threads > MAX_THREADS) { g_last_error = cudaErrorInvalidValue; return; }
    if (run) { fprintf(stderr, "shim_fiber: nested launch\n"); abort(); }
    const ShimThreadState saved = shim_ts;
    pool().launch(blocks, threads, body);
    unsigned char *smem = shim_ts.dyn_smem;
    shim_ts = saved;
    shim_ts.dyn_smem = smem;
}

uint64_t *shim_xchg() { return run->xchg; }

void shim_sync_block() {
    BlockRun &r = *run;
    const unsigned g = r.sync_gen;
    ++r.sync_cnt;
    progress();
    release_if_complete_block();
    while (r.sync_gen == g) yield_block();
}

void shim_sync_warp() {
    BlockRun &r = *run;
    const unsigned w = r.cur >> 5, g = r.wb_gen[w];
    ++r.wb_cnt[w];
    progress();
    release_if_complete_warp(w);
    while (r.wb_gen[w] == g) yield_lane();
}

void shim_sync_mask(unsigned mask) {
    BlockRun &r = *run;
    const unsigned w = r.cur >> 5, g = r.mb_gen[w];
    unsigned n = 0;
    for (unsigned l = 0; l < 32; ++l) {
        const unsigned c = (w << 5) | l;
        if ((mask >> l & 1u) && c < r.nthreads && !r.fibers[c].done) ++n;
    }
    progress();
    if (++r.mb_cnt[w] >= n) { r.mb_cnt[w] = 0; ++r.mb_gen[w]; }
    while (r.mb_gen[w] == g) yield_lane();
}

void shim_named_barrier(int id, unsigned n_threads) {
    BlockRun &r = *run;
    const unsigned g = r.nb_gen[id];
    progress();
    if (++r.nb_cnt[id] >= n_threads) { r.nb_cnt[id] = 0; ++r.nb_gen[id]; }
    while (r.nb_gen[id] == g) yield_block();
}

void shim_yield_block() { yield_block(); }

// ---- mbarrier: arrival count + transaction bytes + phase parity (PTX ISA semantics) --------------------------------------
static void mbar_settle(Mbar &b) {
    if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.init; progress(); }
}
void shim_mbar_init(uint64_t *bar, uint32_t count) {
    Mbar &b = run->mbars[bar];
    b = Mbar{};
    b.init = b.pending = count;
}
void shim_mbar_arrive(uint64_t *bar, uint32_t tx_bytes) {
    Mbar &b = run->mbars.at(bar);
    b.tx += tx_bytes;
    b.pending -= 1;
    progress();
    mbar_settle(b);
}
void shim_bulk_copy(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | bytes) & 15u) {
        fprintf(stderr, "shim_fiber: cp.async.bulk needs 16-byte aligned addresses and sizes\n");
        abort();
    }
    memcpy(dst, src, bytes);
    Mbar &b = run->mbars.at(bar);
    b.tx -= bytes;
    progress();
    mbar_settle(b);
}
void shim_mbar_wait(uint64_t *bar, uint32_t parity) {
    Mbar &b = run->mbars.at(bar);
    while (b.phase == (parity & 1u)) yield_block();
}

// ---- host runtime ---------------------------------------------------------------------------------------------------------------------------
struct ShimStream { int id; };
struct ShimEvent { std::chrono::steady_clock::time_point t; bool recorded = false; };

// ---- "device" memory.  Default: the process heap.  MC_SHIM_SHARED_HEAP=1 (decomposed runs, one process per rank): a
// shared-memory arena per process (bump allocator, nothing is returned before exit), so that a neighbour rank can map it
// and the peer-memory halo of comm.cu / halo_sync.cuh -- stores into the neighbour's arrays, flag words, spins -- runs
// between host processes the way it runs over NVLink.
namespace {
constexpr size_t ARENA_BYTES = (size_t)4 << 30;  // virtual; pages exist once touched
struct Arena {
    char *base = nullptr;
    size_t used = 0;
    char name[64] = {0};
    std::mutex mu;
    std::map<uintptr_t, size_t> allocs;            // base address -> size
    std::map<int, char *> peers;                   // pid -> where that process's arena is mapped here
};
Arena &arena() { static Arena *a = new Arena(); return *a; }
bool shared_heap() { static const bool on = getenv("MC_SHIM_SHARED_HEAP") && atoi(getenv("MC_SHIM_SHARED_HEAP")) != 0; return on; }
void arena_unlink() { if (arena().name[0]) shm_unlink(arena().name); }
bool arena_init() {
    Arena &a = arena();
    if (a.base) return true;
    snprintf(a.name, sizeof(a.name), "/mc_shim_heap_%d", (int)getpid());
    shm_unlink(a.name);
    const int fd = shm_open(a.name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)ARENA_BYTES) != 0) return false;
    void *m = mmap(nullptr, ARENA_BYTES, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return false;
    a.base = static_cast<char *>(m);
    atexit(arena_unlink);
    return true;
}
struct IpcHandle { uint32_t magic; int pid; uint64_t offset; };
CUresult shim_mem_get_address_range(CUdeviceptr *base, size_t *size, CUdeviceptr p) {
    Arena &a = arena();
    std::lock_guard<std::mutex> lk(a.mu);
    auto it = a.allocs.upper_bound((uintptr_t)p);
    if (it == a.allocs.begin()) return 1;
    --it;
    if ((uintptr_t)p >= it->first + it->second) return 1;
    *base = it->first;
    *size = it->second;
    return CUDA_SUCCESS;
}
}  // namespace

cudaError_t shim_malloc(void **p, size_t bytes) {
    if (bytes == 0) bytes = 16;
    bytes = (bytes + 255) & ~(size_t)255;
    void *q = nullptr;
    if (shared_heap()) {
        Arena &a = arena();
        std::lock_guard<std::mutex> lk(a.mu);
        if (!arena_init() || a.used + bytes > ARENA_BYTES) { *p = nullptr; return g_last_error = cudaErrorMemoryAllocation; }
        q = a.base + a.used;
        a.used += bytes;
        a.allocs[(uintptr_t)q] = bytes;
    } else if (posix_memalign(&q, 256, bytes) != 0) {
        *p = nullptr;
        return g_last_error = cudaErrorMemoryAllocation;
    }
    memset(q, 0xFF, bytes);
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) {
    if (!shared_heap()) free(p);  // the arena only grows: short test runs
    return cudaSuccess;
}
cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }

cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *base) {
    if (!shared_heap()) return cudaErrorNotSupported;
    Arena &a = arena();
    std::lock_guard<std::mutex> lk(a.mu);
    if (!a.allocs.count((uintptr_t)base)) return g_last_error = cudaErrorInvalidValue;
    memset(h, 0, sizeof(*h));
    const IpcHandle v{0x4D43u, (int)getpid(), (uint64_t)(static_cast<char *>(base) - a.base)};
    memcpy(h->reserved, &v, sizeof(v));
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
    if (!shared_heap()) return cudaErrorNotSupported;
    IpcHandle v;
    memcpy(&v, h.reserved, sizeof(v));
    if (v.magic != 0x4D43u) return g_last_error = cudaErrorInvalidValue;
    Arena &a = arena();
    std::lock_guard<std::mutex> lk(a.mu);
    char *&map = a.peers[v.pid];
    if (!map) {
        char name[64];
        snprintf(name, sizeof(name), "/mc_shim_heap_%d", v.pid);
        const int fd = shm_open(name, O_RDWR, 0600);
        if (fd < 0) return g_last_error = cudaErrorInvalidValue;
        void *m = mmap(nullptr, ARENA_BYTES, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
        close(fd);
        if (m == MAP_FAILED) return g_last_error = cudaErrorMemoryAllocation;
        map = static_cast<char *>(m);
    }
    *p = map + v.offset;
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }  // mappings of a peer's arena stay until exit
cudaError_t cudaGetDriverEntryPoint(const char *name, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *q) {
    if (shared_heap() && strcmp(name, "cuMemGetAddressRange") == 0) {
        *fn = reinterpret_cast<void *>(&shim_mem_get_address_range);
        if (q) *q = cudaDriverEntryPointSuccess;
        return cudaSuccess;
    }
    *fn = nullptr;
    if (q) *q = cudaDriverEntryPointSymbolNotFound;
    return cudaErrorNotSupported;
}
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind) { if (bytes) memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t) { if (bytes) memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t bytes) { if (bytes) memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t) { if (bytes) memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaSetDevice(int dev) { return dev >= 0 && dev < 8 ? cudaSuccess : (g_last_error = cudaErrorInvalidValue); }  // ranks of a decomposed run name devices 0..7
cudaError_t cudaGetDevice(int *dev) { *dev = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 8; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "host stand-in (tests/cpp/shim_fiber)");
    p->major = 10; p->minor = 0;
    p->multiProcessorCount = 4;
    p->l2CacheSize = 1 << 20;
    p->totalGlobalMem = (size_t)8 << 30;
    p->sharedMemPerBlockOptin = 227 * 1024;
    return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaGetLastError() { const int e = g_last_error; g_last_error = 0; return e; }
cudaError_t cudaPeekAtLastError() { return g_last_error; }
const char *cudaGetErrorString(cudaError_t e) {
    switch (e) {
        case cudaSuccess: return "no error";
        case cudaErrorMemoryAllocation: return "out of memory";
        case cudaErrorInvalidValue: return "invalid argument";
        case cudaErrorNotReady: return "device not ready";
        case 801: return "operation not supported (host stand-in)";
        default: return "unknown error (host stand-in)";
    }
}
cudaError_t cudaStreamCreate(cudaStream_t *st) { *st = new ShimStream{1}; return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *st, unsigned) { *st = new ShimStream{1}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t st) { delete st; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *ev) { *ev = new ShimEvent(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *ev, unsigned) { *ev = new ShimEvent(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t ev) { delete ev; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t ev, cudaStream_t) { ev->t = std::chrono::steady_clock::now(); ev->recorded = true; return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    if (!a->recorded || !b->recorded) return g_last_error = cudaErrorInvalidValue;
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
