// Test infrastructure: the ten NCCL entry points nccl_dyn.cuh resolves with dlopen, over POSIX shared memory, so that
// the decomposed path of the host build (one PROCESS per rank, as on the GPUs) can exchange its blocks on one machine
// without a GPU.  MOLCHANICA_NCCL_LIB points the host build at this file's shared object.
//   unique id   = name of a shared-memory segment; rank 0 creates it, the others attach, rank 0 unlinks it once all have
//   point to point = one byte pipe (ring buffer) per ordered pair of ranks: FIFO per pair, exactly NCCL's matching rule.
//                 Operations posted between ncclGroupStart / ncclGroupEnd are progressed together, a few bytes of each at a
//                 time, so that no ordering of sends and receives on the two sides can deadlock on a full pipe.
//   collectives = every rank writes its contribution into a slot of a shared buffer, barrier, every rank reads, barrier.
// "Streams" are ignored: the host stand-in of the CUDA runtime is synchronous.
#include <fcntl.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <vector>

#define MC_NCCL_STANDIN_NO_RUNTIME 1
#include "../shim_fiber/nccl.h"

namespace {

constexpr int MAX_RANKS = 8;
constexpr size_t PIPE_BYTES = 1 << 20;
constexpr size_t COLL_BYTES = (size_t)256 << 20;  // sparse until touched

struct Pipe {
    std::atomic<uint64_t> head, tail;  // bytes written / read so far
    char pad[48];
    unsigned char data[PIPE_BYTES];
};

struct Segment {
    std::atomic<uint32_t> attached, bar_count, bar_gen;
    uint32_t n_ranks;
    char pad[48];
    Pipe pipes[MAX_RANKS * MAX_RANKS];  // [src * MAX_RANKS + dst]
    unsigned char coll[COLL_BYTES];
};

struct Op { bool send; unsigned char *p; size_t bytes, done; int peer; };

}  // namespace

struct ncclComm {
    Segment *seg;
    int rank, n;
    std::vector<Op> pending;
};

namespace {

thread_local int group_depth = 0;
thread_local ncclComm *group_comm = nullptr;

size_t type_bytes(ncclDataType_t t) {
    switch (t) {
        case ncclInt8: case ncclUint8: return 1;
        case ncclFloat16: return 2;
        case ncclInt32: case ncclUint32: case ncclFloat32: return 4;
        default: return 8;
    }
}

void barrier(ncclComm *c) {
    Segment *s = c->seg;
    const uint32_t g = s->bar_gen.load();
    if (s->bar_count.fetch_add(1) + 1 == (uint32_t)c->n) {
        s->bar_count.store(0);
        s->bar_gen.fetch_add(1);
    } else {
        while (s->bar_gen.load() == g) sched_yield();
    }
}

// moves as many bytes of one operation as the pipe allows right now; true when the operation is complete
bool progress(ncclComm *c, Op &o) {
    if (o.done == o.bytes) return true;
    Pipe &p = o.send ? c->seg->pipes[c->rank * MAX_RANKS + o.peer] : c->seg->pipes[o.peer * MAX_RANKS + c->rank];
    const uint64_t head = p.head.load(std::memory_order_acquire), tail = p.tail.load(std::memory_order_acquire);
    if (o.send) {
        size_t room = PIPE_BYTES - (size_t)(head - tail), n = o.bytes - o.done < room ? o.bytes - o.done : room;
        for (size_t k = 0; k < n;) {
            const size_t at = (size_t)((head + k) % PIPE_BYTES), run = n - k < PIPE_BYTES - at ? n - k : PIPE_BYTES - at;
            memcpy(p.data + at, o.p + o.done + k, run);
            k += run;
        }
        p.head.store(head + n, std::memory_order_release);
        o.done += n;
    } else {
        size_t avail = (size_t)(head - tail), n = o.bytes - o.done < avail ? o.bytes - o.done : avail;
        for (size_t k = 0; k < n;) {
            const size_t at = (size_t)((tail + k) % PIPE_BYTES), run = n - k < PIPE_BYTES - at ? n - k : PIPE_BYTES - at;
            memcpy(o.p + o.done + k, p.data + at, run);
            k += run;
        }
        p.tail.store(tail + n, std::memory_order_release);
        o.done += n;
    }
    return o.done == o.bytes;
}

// Within one pipe the operations must complete in posting order (a byte pipe has no message boundaries): only the first
// unfinished send and the first unfinished receive of every peer are progressed.
void drain(ncclComm *c) {
    for (;;) {
        bool all = true;
        bool busy_send[MAX_RANKS] = {false}, busy_recv[MAX_RANKS] = {false};
        for (Op &o : c->pending) {
            bool *busy = o.send ? busy_send : busy_recv;
            if (o.done == o.bytes) continue;
            if (!busy[o.peer]) {
                if (!progress(c, o)) busy[o.peer] = true;
            }
            if (o.done != o.bytes) all = false;
        }
        if (all) break;
        sched_yield();
    }
    c->pending.clear();
}

ncclResult_t post(ncclComm *c, bool send, const void *buf, size_t count, ncclDataType_t t, int peer) {
    if (peer < 0 || peer >= c->n) return ncclInvalidArgument;
    c->pending.push_back(Op{send, static_cast<unsigned char *>(const_cast<void *>(buf)), count * type_bytes(t), 0, peer});
    if (group_depth == 0) drain(c);
    else group_comm = c;
    return ncclSuccess;
}

template <typename T>
void reduce(T *out, const unsigned char *slots, size_t count, size_t stride, int n, ncclRedOp_t op) {
    for (size_t i = 0; i < count; ++i) {
        T acc = reinterpret_cast<const T *>(slots)[i];
        for (int r = 1; r < n; ++r) {
            const T v = reinterpret_cast<const T *>(slots + (size_t)r * stride)[i];
            acc = op == ncclSum ? acc + v : op == ncclMax ? (v > acc ? v : acc) : op == ncclMin ? (v < acc ? v : acc) : acc * v;
        }
        out[i] = acc;
    }
}

}  // namespace

extern "C" {

ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
    memset(id, 0, sizeof(*id));
    struct timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    snprintf(id->internal, sizeof(id->internal), "/mc_nccl_standin_%d_%ld_%ld", (int)getpid(), (long)ts.tv_sec, (long)ts.tv_nsec);
    return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t *out, int n, ncclUniqueId id, int rank) {
    if (n < 1 || n > MAX_RANKS || rank < 0 || rank >= n) return ncclInvalidArgument;
    int fd = -1;
    if (rank == 0) {
        fd = shm_open(id.internal, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)sizeof(Segment)) != 0) return ncclSystemError;
    } else {
        for (int tries = 0; tries < 60000 && fd < 0; ++tries) {
            fd = shm_open(id.internal, O_RDWR, 0600);
            if (fd < 0) usleep(1000);
        }
        if (fd < 0) return ncclSystemError;
        // the creator sizes the segment before anybody maps it
        for (int tries = 0; tries < 60000; ++tries) {
            off_t len = lseek(fd, 0, SEEK_END);
            if (len >= (off_t)sizeof(Segment)) break;
            usleep(1000);
        }
    }
    void *m = mmap(nullptr, sizeof(Segment), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return ncclSystemError;
    ncclComm *c = new ncclComm{static_cast<Segment *>(m), rank, n, {}};
    if (rank == 0) c->seg->n_ranks = (uint32_t)n;  // a fresh segment is zero-filled: counters and pipes start empty
    c->seg->attached.fetch_add(1);
    while (c->seg->attached.load() < (uint32_t)n) usleep(200);
    barrier(c);
    if (rank == 0) shm_unlink(id.internal);  // everybody holds a mapping: the name is no longer needed
    *out = c;
    return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t c) {
    if (!c) return ncclSuccess;
    munmap(c->seg, sizeof(Segment));
    delete c;
    return ncclSuccess;
}

const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : r == ncclSystemError ? "system error (shared memory)" : "invalid argument"; }

ncclResult_t ncclGroupStart() { ++group_depth; return ncclSuccess; }
ncclResult_t ncclGroupEnd() {
    if (--group_depth == 0 && group_comm) { drain(group_comm); group_comm = nullptr; }
    return ncclSuccess;
}

ncclResult_t ncclSend(const void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, void *) { return post(c, true, buf, count, t, peer); }
ncclResult_t ncclRecv(void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, void *) { return post(c, false, buf, count, t, peer); }

ncclResult_t ncclAllGather(const void *send, void *recv, size_t count, ncclDataType_t t, ncclComm_t c, void *) {
    const size_t bytes = count * type_bytes(t);
    if (bytes * c->n > COLL_BYTES) return ncclInvalidArgument;
    memcpy(c->seg->coll + (size_t)c->rank * bytes, send, bytes);
    barrier(c);
    memcpy(recv, c->seg->coll, bytes * c->n);
    barrier(c);
    return ncclSuccess;
}

ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t c, void *) {
    const size_t bytes = count * type_bytes(t);
    if (bytes * c->n > COLL_BYTES) return ncclInvalidArgument;
    memcpy(c->seg->coll + (size_t)c->rank * bytes, send, bytes);
    barrier(c);
    switch (t) {
        case ncclInt32: reduce(static_cast<int32_t *>(recv), c->seg->coll, count, bytes, c->n, op); break;
        case ncclUint32: reduce(static_cast<uint32_t *>(recv), c->seg->coll, count, bytes, c->n, op); break;
        case ncclFloat32: reduce(static_cast<float *>(recv), c->seg->coll, count, bytes, c->n, op); break;
        case ncclFloat64: reduce(static_cast<double *>(recv), c->seg->coll, count, bytes, c->n, op); break;
        default: return ncclInvalidArgument;
    }
    barrier(c);
    return ncclSuccess;
}
}
