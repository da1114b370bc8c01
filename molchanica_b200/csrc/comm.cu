// comm.cu -- spatial domain decomposition over the GPUs of one box: slabs of whole cell
// layers along z, one ghost layer on each side, one process per GPU.
//
// Layout on a rank (cell order, z slowest):   [ ghost layer from prev | owned layers | ghost layer from next ]
// Because the cell sort is z-major, each of the three parts -- and the first / last OWNED layer that
// the neighbours need as their ghosts -- is one contiguous block of the position array, ordered
// identically on the owner and on the ghost holder (see comm_rebuild).  A ghost refresh is therefore a
// plain block copy with no pack / unpack kernels.  Full (not half) neighbour lists mean every rank
// computes the forces of its own atoms only: there is no reverse (force) communication.
//
// Per-step ghost refresh, two implementations:
//   fused (default)  the neighbours' position arrays and flag words are mapped into this process with
//                    cudaIpc at the first build; kick_drift stores its boundary layers straight into the
//                    neighbours' ghost blocks over NVLink and raises epoch flags, the pair kernel's boundary
//                    blocks wait on them (halo_sync.cuh, integrate.cu, pair_force.cu).  No call here per step:
//                    comm_step_descriptors only fills the descriptors the two kernels take.
//   NCCL             comm_halo_positions: ncclSend / ncclRecv of the four blocks between the two kernels
//                    (option halo_fused = 0, or the mapping failed on some rank).
//
// Positions are kept in the GLOBAL frame and the minimum image is applied to pair differences,
// exactly as on one GPU (ghosts are never shifted by +-L, which would cost ~1e-5 A of fp32
// resolution at |z| ~ L), so a decomposed run lists and masks the same pairs as the single-GPU run.
//
// Rebuild: the first build all-gathers every rank's atoms in fixed-capacity blocks (padding id -1) and
// keeps what falls into the rank's layer range [kz0-1, kz1] by a filter key + the same radix sort /
// reorder as the single-GPU path.  Later builds exchange only the last two / first two owned layers
// with the ring neighbours (block sizes are known on both sides from the layout record every build
// all-gathers), laid out [from prev | own | from next] so that the stable sort orders a layer the same
// way on both sides.  The schedule is derived from all-gathered numbers, identical on every rank.
#include "nccl_dyn.cuh"

#include <cuda.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

#include "engine.cuh"
#include "integrate.cuh"
#include "neighbor.cuh"

struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, n = 1;
    int ncz = 0, kz0 = 0, kz1 = 0, nl = 0;  // global layers, owned range [kz0, kz1), owned layer count
    size_t cap = 0;                         // atoms per rank in the all-gather blocks
    size_t local_cap = 0;                   // capacity of the local (owned + ghost) arrays
    DevBuf<float4> s_xyzq, s_vel, g_xyzq, g_vel;
    DevBuf<int2> s_meta, g_meta;            // {original id (-1 = padding), type | flags << 16}
    DevBuf<uint32_t> d_layer;               // 6 slot offsets read back after every rebuild
    DevBuf<double> d_red;
    DevBuf<int> d_flags_all;                // every rank's flag words (+ this rank's own two behind them)
    // slot offsets of the current build
    uint32_t o_gp = 0, o_own = 0, o_first_end = 0, o_last_begin = 0, o_own_end = 0, o_end = 0;

    // ---- peer-memory halo (halo_sync.cuh): the neighbours' position arrays and flag words mapped
    // into this process with cudaIpc, so that kick_drift stores ghosts straight into them over NVLink
    struct PeerMap {
        void *opened[3] = {nullptr, nullptr, nullptr};  // what cudaIpcOpenMemHandle returned (to close)
        float4 *xyzq[2] = {nullptr, nullptr};
        uint32_t *flags = nullptr;
    };
    bool peer_tried = false, peer_ok = false;
    std::string peer_why;                   // why the fused halo is off (when it is)
    DevBuf<uint32_t> flags;                 // this rank's flag words (exported)
    DevBuf<uint8_t> d_exp, d_exp_all;       // handle exchange staging
    DevBuf<uint32_t> d_layer_all;           // every rank's slot offsets of the current build
    uint32_t *h_layer_all = nullptr;        // pinned
    const void *exported[2] = {nullptr, nullptr};
    PeerMap prev_map, next_map;
    bool have_table = false;                // h_layer_all describes the build the local arrays are in
    bool dd_migrate = true;                 // rebuilds exchange boundary layers with the two neighbours only
    bool preconnected = false;              // the ring's point-to-point channels have been set up (first build)
    DevBuf<float4> pre_buf;                 // scratch of that warm-up exchange
    int interval = 10;                      // adaptive rebuild interval (option rebuild_every = 0)
    float vmax0 = 0.f;                      // largest speed handed to mc_set_atoms (every rank sees the whole array: same number)
    bool first_interval_pending = false;    // the first interval is still the default: comm_first_interval may shorten it
    double last_disp_frac = 0.0;            // largest displacement of the last interval / (skin/2)
    uint32_t epoch = 1;                     // the first step runs at epoch 2: its acks are real signals
    float4 *to_prev = nullptr, *to_next = nullptr;  // ghost blocks of the current build on the neighbours
};

#define MC_NCCL(ctx, call)                                                                              \
    do {                                                                                                \
        ncclResult_t r_ = (call);                                                                       \
        if (r_ != ncclSuccess) {                                                                        \
            (ctx)->err = std::string(#call) + ": " + nccl_api().GetErrorString(r_);                            \
            return MC_E_COMM;                                                                           \
        }                                                                                               \
    } while (0)

#define MC_CUDAC(ctx, call)                                                                             \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            return MC_E_CUDA;                                                                           \
        }                                                                                               \
    } while (0)

namespace {

// owned block -> fixed-capacity send buffers, padding marked with id -1
__global__ void dd_pack_kernel(int n_own, int cap, const float4 *__restrict__ xyzq, const float4 *__restrict__ vel,
                               const int *__restrict__ orig, const uint16_t *__restrict__ type,
                               const uint8_t *__restrict__ flags, float4 *__restrict__ s_xyzq, float4 *__restrict__ s_vel,
                               int2 *__restrict__ s_meta) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cap) return;
    if (k < n_own) {
        s_xyzq[k] = xyzq[k];
        s_vel[k] = vel[k];
        s_meta[k] = make_int2(orig[k], (int)type[k] | ((int)(flags[k] & (uint8_t)~MC_FLAG_INTERIOR) << 16));
    } else {
        s_meta[k] = make_int2(-1, 0);
    }
}

// gathered atoms -> sort key: local cell id for atoms inside this rank's layer range, a sentinel
// (== number of local cells, sorts last) for everything else.  Positions are wrapped into the box.
__global__ void dd_key_kernel(int n_all, float4 *__restrict__ g_xyzq, const int2 *__restrict__ g_meta,
                              const GridParams *__restrict__ gp, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_all) return;
    const GridParams g = *gp;
    vals[i] = (uint32_t)i;
    if (g_meta[i].x < 0) { keys[i] = (uint32_t)g.ncell << g.sub_bits; return; }
    float4 p = g_xyzq[i];
    float c[3] = {p.x, p.y, p.z};
    float fr[3];
    int k[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        c[a] -= floorf((c[a] - g.lo[a]) * g.inv_ext[a]) * g.ext[a];
        if (c[a] < g.lo[a]) c[a] += g.ext[a];
        if (c[a] >= g.lo[a] + g.ext[a]) c[a] -= g.ext[a];
        const int nca = a == 2 ? g.ncz_global : g.nc[a];
        const float u = (c[a] - g.lo[a]) * g.inv_cw[a];
        const int kk = (int)floorf(u);
        k[a] = min(max(kk, 0), nca - 1);
        fr[a] = u - (float)k[a];
    }
    g_xyzq[i] = make_float4(c[0], c[1], c[2], p.w);
    const int l = (k[2] - g.kz_off + g.ncz_global) % g.ncz_global;  // local layer of this global layer
    const uint32_t cell = l < g.nc[2] ? (uint32_t)((l * g.nc[1] + k[1]) * g.nc[0] + k[0]) : (uint32_t)g.ncell;
    keys[i] = g.sub_bits ? (cell << g.sub_bits) | (l < g.nc[2] ? mc_subcell_code(fr[0], fr[1], fr[2]) : 0u) : cell;
}

// sorted (key, index) -> local cell-ordered arrays + cell_start; entries with the sentinel key are dropped
__global__ void dd_reorder_kernel(int n_all, const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ svals,
                                  const GridParams *__restrict__ gp, const float4 *__restrict__ g_xyzq,
                                  const float4 *__restrict__ g_vel, const int2 *__restrict__ g_meta, int mark_interior,
                                  float4 *__restrict__ xyzq, float4 *__restrict__ xref, float4 *__restrict__ vel,
                                  uint16_t *__restrict__ type, uint8_t *__restrict__ flags, int *__restrict__ orig,
                                  int *__restrict__ slot_of_orig, uint32_t *__restrict__ cell_start, uint32_t local_cap) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n_all) return;
    const int ncell = gp->ncell, sb = gp->sub_bits;
    const int prev = k == 0 ? -1 : (int)(skeys[k - 1] >> sb);
    const int cur = k == n_all ? ncell + 1 : (int)(skeys[k] >> sb);
    // cell_start[0 .. ncell]: cell_start[ncell] = number of local atoms (first sentinel entry)
    for (int c = prev + 1; c <= min(cur, ncell); ++c) cell_start[c] = (uint32_t)k;
    if (k == n_all || cur >= ncell || (uint32_t)k >= local_cap) return;
    const uint32_t src = svals[k];
    const float4 p = g_xyzq[src];
    xyzq[k] = p;
    xref[k] = p;
    vel[k] = g_vel[src];
    const int2 m = g_meta[src];
    type[k] = (uint16_t)(m.y & 0xffff);
    uint8_t fl = (uint8_t)((m.y >> 16) & 0xff);
    if (mark_interior) {
        const int nc0 = gp->nc[0], nc1 = gp->nc[1], nc2 = gp->ncz_global;
        const int c0 = cur % nc0, c1 = (cur / nc0) % nc1, c2 = (cur / (nc0 * nc1) + gp->kz_off) % nc2;
        if (nc0 >= 3 && nc1 >= 3 && nc2 >= 3 && c0 >= 1 && c0 <= nc0 - 2 && c1 >= 1 && c1 <= nc1 - 2 && c2 >= 1 &&
            c2 <= nc2 - 2)
            fl |= MC_FLAG_INTERIOR;
    }
    flags[k] = fl;
    orig[k] = m.x;
    slot_of_orig[m.x] = k;
}

#define MC_DD_TABLE_WORDS 16
// Per-rank record of a build, all-gathered so that every rank knows its neighbours' layout:
//   [0] first local slot  [1] first owned slot  [2] end of the first owned layer  [3] begin of the last owned layer
//   [4] end of the owned slots  [5] end of the local slots  [6] which position array holds the build
//   [7] largest squared displacement since the PREVIOUS build (float bits; read and reset here)
//   [8] end of the second owned layer  [9] begin of the second-to-last owned layer
__global__ void dd_layer_offsets_kernel(const uint32_t *__restrict__ cell_start, int plane, int nlayers, uint32_t cur,
                                        int *__restrict__ max_disp2_word, uint32_t *__restrict__ out) {
    if (threadIdx.x != 0) return;
    const int nl = nlayers - 2;  // owned layers
    out[0] = cell_start[0];
    out[1] = cell_start[plane];                  // first owned slot
    out[2] = cell_start[2 * plane];              // end of the first owned layer
    out[3] = cell_start[(nlayers - 2) * plane];  // begin of the last owned layer
    out[4] = cell_start[(nlayers - 1) * plane];  // end of the owned slots
    out[5] = cell_start[nlayers * plane];        // end of the local slots
    out[6] = cur;
    out[7] = (uint32_t)*max_disp2_word;
    *max_disp2_word = 0;
    out[8] = cell_start[(nl >= 2 ? 3 : 2) * plane];
    out[9] = cell_start[(nl >= 2 ? nlayers - 3 : nlayers - 2) * plane];
    for (int k = 10; k < MC_DD_TABLE_WORDS; ++k) out[k] = 0;
}

}  // namespace

extern "C" int mc_comm_unique_id(uint8_t id[128]) {
    if (!id) return MC_E_INVALID;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    if (!nccl_api().ok || nccl_api().GetUniqueId(&u) != ncclSuccess) return MC_E_COMM;
    memcpy(id, &u, 128);
    return MC_OK;
}

extern "C" int mc_comm_init(mc_ctx *c, const uint8_t id[128], int rank, int n_ranks) {
    if (!c || !id) return MC_E_INVALID;
    if (n_ranks < 2 || rank < 0 || rank >= n_ranks) { c->err = "mc_comm_init: need n_ranks >= 2 and 0 <= rank < n_ranks"; return MC_E_INVALID; }
    if (c->n != 0) { c->err = "mc_comm_init: must precede mc_set_atoms"; return MC_E_INVALID; }
    cudaSetDevice(c->device);
    if (!nccl_api().ok) { c->err = "mc_comm_init: " + nccl_api().err; return MC_E_COMM; }
    g_devbuf_roomy = true;  // buffers that follow the atoms a rank holds: head-room from the first allocation on (engine.cuh)
    CommState *cs = new CommState();
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclResult_t r = nccl_api().CommInitRank(&cs->comm, n_ranks, u, rank);
    if (r != ncclSuccess) {
        c->err = std::string("ncclCommInitRank: ") + nccl_api().GetErrorString(r);
        delete cs;
        return MC_E_COMM;
    }
    cs->rank = rank;
    cs->n = n_ranks;
    c->comm = cs;
    c->comm_active = true;
    return MC_OK;
}

extern "C" int mc_comm_counts(mc_ctx *c, int64_t *n_owned, int64_t *n_ghost) {
    if (!c) return MC_E_INVALID;
    if (n_owned) *n_owned = c->n_rows;
    if (n_ghost) *n_ghost = c->n - c->n_rows;
    return MC_OK;
}

extern "C" int mc_get_positions_global(mc_ctx *c, mc_float4 *out) { return mc_get_positions(c, out); }
extern "C" int mc_get_forces_global(mc_ctx *c, mc_float4 *out) { return mc_get_forces(c, out); }

static void peer_close(CommState *cs) {
    CommState::PeerMap *maps[2] = {&cs->prev_map, &cs->next_map};
    for (int m = 0; m < 2; ++m) {
        for (int k = 0; k < 3; ++k) {
            void *p = maps[m]->opened[k];
            if (!p) continue;
            bool dup = false;  // prev == next (two ranks) or several arrays in one allocation share a mapping
            for (int m2 = 0; m2 <= m && !dup; ++m2)
                for (int k2 = 0; k2 < (m2 == m ? k : 3) && !dup; ++k2) dup = maps[m2]->opened[k2] == p;
            if (!dup) cudaIpcCloseMemHandle(p);
        }
        *maps[m] = CommState::PeerMap();
    }
    cs->peer_ok = false;
    cs->to_prev = cs->to_next = nullptr;
}

void comm_destroy(mc_ctx *c) {
    if (!c->comm) return;
    CommState *cs = c->comm;
    peer_close(cs);
    cs->flags.release(); cs->d_exp.release(); cs->d_exp_all.release(); cs->d_layer_all.release(); cs->pre_buf.release();
    if (cs->h_layer_all) cudaFreeHost(cs->h_layer_all);
    if (cs->comm) nccl_api().CommDestroy(cs->comm);
    cs->s_xyzq.release(); cs->s_vel.release(); cs->g_xyzq.release(); cs->g_vel.release();
    cs->s_meta.release(); cs->g_meta.release(); cs->d_layer.release(); cs->d_red.release(); cs->d_flags_all.release();
    delete cs;
    c->comm = nullptr;
    c->comm_active = false;
}

// Global system on every rank -> this rank starts with an index block of it; the first rebuild
// redistributes by position.
int comm_set_atoms(mc_ctx *c, int64_t n, const mc_float4 *xyzq, const uint16_t *type, const mc_float4 *vel,
                   const uint8_t *flags) {
    CommState *cs = c->comm;
    if (!c->periodic) { c->err = "domain decomposition needs a periodic box"; return MC_E_INVALID; }
    const int64_t lo = n * cs->rank / cs->n, hi = n * (cs->rank + 1) / cs->n, m = hi - lo;
    cs->cap = (size_t)(n / cs->n + n / (2 * cs->n) + 4096);        // 1.5 x the mean + slack
    cs->local_cap = std::min<size_t>((size_t)n + 64, 3 * cs->cap);  // owned + two ghost layers
    std::vector<int> ids((size_t)m);
    for (int64_t k = 0; k < m; ++k) ids[(size_t)k] = (int)(lo + k);
    c->n_global = n;
    // allocate for the largest local population, then upload the initial block
    cudaError_t e = c->alloc_atoms(cs->local_cap);
    if (e != cudaSuccess) { c->err = std::string("alloc_atoms: ") + cudaGetErrorString(e); return MC_E_CUDA; }
    int rc = engine_upload_local(c, m, xyzq + lo, type ? type + lo : nullptr, vel ? vel + lo : nullptr,
                                 flags ? flags + lo : nullptr, ids.data(), cs->local_cap);
    if (rc != MC_OK) return rc;
    c->n = m;
    c->n_rows = m;
    c->row0 = 0;
    cs->have_table = false;  // the atoms are index blocks again: the next build is an all-gather
    cs->interval = 10;
    // the fastest atom of the system as handed over (for the first interval, see comm_first_interval)
    double v2 = 0.0;
    if (vel)
        for (int64_t k = 0; k < n; ++k)
            if (vel[k].w > 0.f) v2 = std::max(v2, (double)vel[k].x * vel[k].x + (double)vel[k].y * vel[k].y + (double)vel[k].z * vel[k].z);
    cs->vmax0 = (float)std::sqrt(v2);
    cs->first_interval_pending = true;
    MC_CUDAC(c, cudaMemset(c->slot_of_orig.p, 0xff, sizeof(int) * (size_t)n));
    return MC_OK;
}

// The adaptive schedule learns the interval from the displacements of the interval before; the very first one has nothing to
// learn from and used to be 10 steps whatever the system -- too long for a hot fluid under a thin skin (900 K argon, 0.6 A:
// the first interval overshot skin/2 and was reported as MC_W_STALE_LIST; found by tools/dd_fuzz_host.py).  With the time step
// of the first mc_step known: the fastest atom, moving ballistically, may cover half of skin/2 -- between 4 and 10 steps.
// Every rank computes it from the same numbers.
void comm_first_interval(mc_ctx *c, float dt) {
    CommState *cs = c->comm;
    if (!cs || !cs->first_interval_pending) return;
    cs->first_interval_pending = false;
    const double per_step = (double)cs->vmax0 * std::fabs((double)dt);
    if (per_step <= 0.0 || c->skin <= 0.f) return;
    const double k = std::floor(0.5 * (0.5 * (double)c->skin) / per_step);
    cs->interval = (int)std::max(4.0, std::min(10.0, k));
}

// The decomposition arithmetic, shared by the device path below and by mc_dd_plan (host only, no GPU):
// global cell grid of the periodic box and the z layers [kz0, kz1) a rank owns.
static int dd_plan_grid(const float ext[3], float r_list, int rank, int n_ranks, int ncg[3], int *kz0, int *kz1,
                        std::string *err) {
    const double cw_min = (double)r_list * 1.001 + 1e-3;
    for (int a = 0; a < 3; ++a) {
        if (2.0f * r_list > ext[a]) { *err = "cutoff + skin exceeds half the periodic box"; return MC_E_INVALID; }
        int m = (int)std::floor((double)ext[a] / cw_min);
        m = std::max(1, std::min(m, 1024));
        if (a == 2 && m >= 2 * n_ranks) m -= m % n_ranks;  // equal layer counts per rank when the box allows
        ncg[a] = m;
    }
    if (ncg[2] < 2 * n_ranks) {
        *err = "domain decomposition: the box holds " + std::to_string(ncg[2]) + " cell layers along z, need >= 2 per rank";
        return MC_E_INVALID;
    }
    *kz0 = (int)((long long)ncg[2] * rank / n_ranks);
    *kz1 = (int)((long long)ncg[2] * (rank + 1) / n_ranks);
    return MC_OK;
}

extern "C" int mc_dd_plan(const float box_ext[3], float r_list, int rank, int n_ranks, int32_t out[8]) {
    if (!box_ext || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return MC_E_INVALID;
    int ncg[3], kz0, kz1;
    std::string err;
    int rc = dd_plan_grid(box_ext, r_list, rank, n_ranks, ncg, &kz0, &kz1, &err);
    if (rc != MC_OK) return rc;
    out[0] = ncg[0]; out[1] = ncg[1]; out[2] = ncg[2];
    out[3] = kz0; out[4] = kz1;
    out[5] = (kz0 - 1 + ncg[2]) % ncg[2];     // ghost layer received from the previous rank
    out[6] = kz1 % ncg[2];                    // ghost layer received from the next rank
    out[7] = (rank + 1) % n_ranks;            // next rank (prev = (rank + n - 1) % n)
    return MC_OK;
}

static int dd_setup_grid(mc_ctx *c) {
    CommState *cs = c->comm;
    const float r_list = std::max(c->rc_lj, c->rc_q) + c->skin;
    GridParams g;
    int ncg[3];
    int rc = dd_plan_grid(c->ext, r_list, cs->rank, cs->n, ncg, &cs->kz0, &cs->kz1, &c->err);
    if (rc != MC_OK) return rc;
    for (int a = 0; a < 3; ++a) {
        g.lo[a] = c->lo[a];
        g.ext[a] = c->ext[a];
        g.inv_ext[a] = 1.0f / c->ext[a];
        g.inv_cw[a] = (float)((double)ncg[a] / (double)c->ext[a]);
    }
    cs->ncz = ncg[2];
    cs->nl = cs->kz1 - cs->kz0;
    g.nc[0] = ncg[0]; g.nc[1] = ncg[1]; g.nc[2] = cs->nl + 2;
    g.ncell = g.nc[0] * g.nc[1] * g.nc[2];
    g.periodic = 1;
    g.z_ring = 0;
    g.kz_off = (cs->kz0 - 1 + ncg[2]) % ncg[2];
    g.ncz_global = ncg[2];
    g.row_l0 = 1;
    g.row_l1 = cs->nl + 1;
    g.sub_bits = (c->subcell_sort && (long long)g.ncell + 1 <= (1ll << (32 - MC_SUB_BITS - 1))) ? MC_SUB_BITS : 0;
    c->h_grid = g;
    c->ncell_cap = (size_t)g.ncell;
    MC_CUDAC(c, c->grid.ensure(1));
    MC_CUDAC(c, cudaMemcpyAsync(c->grid.p, &c->h_grid, sizeof(GridParams), cudaMemcpyHostToDevice, c->st));
    MC_CUDAC(c, c->cell_start.ensure(c->ncell_cap + 2));
    int bits = 1;
    while (((size_t)1 << bits) < c->ncell_cap + 1) ++bits;  // + sentinel key
    c->key_bits = bits + g.sub_bits;
    c->grid_dirty = false;
    return MC_OK;
}

// ---- peer-memory halo set-up ---------------------------------------------------------------------
// Every rank exports its two position arrays and its flag words as cudaIpc handles (+ the offset of
// the pointer inside its allocation), the handles travel through one ncclAllGather, and each rank
// maps those of its two ring neighbours.  Any failure (no P2P path, IPC refused by the container)
// leaves the NCCL send / recv halo in place -- on all ranks, the outcome is agreed with an all-reduce.
struct PeerExport {
    cudaIpcMemHandle_t h[3];
    uint64_t off[3];
    uint64_t base[3];  // exporter-side base addresses: arrays that share an allocation are mapped once
    uint64_t pad[2];
};
static_assert(sizeof(PeerExport) == 256, "PeerExport is one 256-byte record");

static bool alloc_base(const void *p, const void **base) {
    typedef CUresult (*range_fn)(CUdeviceptr *, size_t *, CUdeviceptr);
    static range_fn fn = nullptr;
    if (!fn) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return false;
        fn = reinterpret_cast<range_fn>(f);
    }
    CUdeviceptr b = 0;
    size_t sz = 0;
    if (fn(&b, &sz, (CUdeviceptr)(uintptr_t)p) != CUDA_SUCCESS) return false;
    *base = reinterpret_cast<const void *>((uintptr_t)b);
    return true;
}

static bool peer_open(const PeerExport &e, CommState::PeerMap *m, std::string *why) {
    char *mapped[3] = {nullptr, nullptr, nullptr};
    for (int k = 0; k < 3; ++k) {
        for (int k2 = 0; k2 < k; ++k2)
            if (e.base[k] == e.base[k2]) mapped[k] = mapped[k2];
        if (!mapped[k]) {
            void *p = nullptr;
            cudaError_t err = cudaIpcOpenMemHandle(&p, e.h[k], cudaIpcMemLazyEnablePeerAccess);
            if (err != cudaSuccess) {
                cudaGetLastError();
                *why = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(err);
                return false;
            }
            mapped[k] = static_cast<char *>(p);
        }
        m->opened[k] = mapped[k];
    }
    m->xyzq[0] = reinterpret_cast<float4 *>(mapped[0] + e.off[0]);
    m->xyzq[1] = reinterpret_cast<float4 *>(mapped[1] + e.off[1]);
    m->flags = reinterpret_cast<uint32_t *>(mapped[2] + e.off[2]);
    return true;
}

static int peer_setup(mc_ctx *c) {
    CommState *cs = c->comm;
    cs->peer_tried = true;
    peer_close(cs);
    const int prev = (cs->rank + cs->n - 1) % cs->n, next = (cs->rank + 1) % cs->n;
    bool ok = true;
    PeerExport mine;
    memset(&mine, 0, sizeof(mine));
    MC_CUDAC(c, cs->flags.ensure(MC_HALO_FLAG_WORDS));
    MC_CUDAC(c, cudaMemsetAsync(cs->flags.p, 0, MC_HALO_FLAG_WORDS * sizeof(uint32_t), c->st));
    cs->epoch = 1;
    const void *ptrs[3] = {c->xyzq[0].p, c->xyzq[1].p, cs->flags.p};
    for (int k = 0; k < 3 && ok; ++k) {
        const void *base = nullptr;
        if (!alloc_base(ptrs[k], &base)) { ok = false; cs->peer_why = "cuMemGetAddressRange unavailable"; break; }
        cudaError_t e = cudaIpcGetMemHandle(&mine.h[k], const_cast<void *>(base));
        if (e != cudaSuccess) { cudaGetLastError(); ok = false; cs->peer_why = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); break; }
        mine.off[k] = (uint64_t)((const char *)ptrs[k] - (const char *)base);
        mine.base[k] = (uint64_t)(uintptr_t)base;
    }
    cs->exported[0] = c->xyzq[0].p;
    cs->exported[1] = c->xyzq[1].p;
    MC_CUDAC(c, cs->d_exp.ensure(sizeof(PeerExport)));
    MC_CUDAC(c, cs->d_exp_all.ensure(sizeof(PeerExport) * (size_t)cs->n));
    MC_CUDAC(c, cudaMemcpyAsync(cs->d_exp.p, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->st));
    MC_NCCL(c, nccl_api().AllGather(cs->d_exp.p, cs->d_exp_all.p, sizeof(PeerExport), ncclUint8, cs->comm, c->st));
    std::vector<PeerExport> all((size_t)cs->n);
    MC_CUDAC(c, cudaMemcpyAsync(all.data(), cs->d_exp_all.p, sizeof(PeerExport) * (size_t)cs->n, cudaMemcpyDeviceToHost, c->st));
    MC_CUDAC(c, cudaStreamSynchronize(c->st));
    if (ok) ok = peer_open(all[(size_t)prev], &cs->prev_map, &cs->peer_why);
    if (ok) {
        if (next == prev) cs->next_map = cs->prev_map;
        else ok = peer_open(all[(size_t)next], &cs->next_map, &cs->peer_why);
    }
    bool bad = !ok;
    int rc = comm_agree_flag(c, &bad);  // max over ranks: one failure switches every rank to NCCL
    if (rc != MC_OK) return rc;
    if (bad) {
        if (ok) cs->peer_why = "another rank could not map its neighbours";
        peer_close(cs);
    } else {
        cs->peer_ok = true;
        cs->peer_why.clear();
    }
    return MC_OK;
}

bool comm_peer_direct(const mc_ctx *c) { return c->comm && c->comm->peer_ok; }

extern "C" int mc_comm_schedule(mc_ctx *c, int *interval, double *last_disp_frac) {
    if (!c || !c->comm) return MC_E_INVALID;
    if (interval) *interval = comm_interval(c);
    if (last_disp_frac) *last_disp_frac = c->comm->last_disp_frac;
    return MC_OK;
}

void comm_set_migrate(mc_ctx *c, bool on) { if (c->comm) c->comm->dd_migrate = on; }

// Steps between two builds of a decomposed run: the caller's fixed schedule, else (fused halo, whose kick_drift
// tracks the largest displacement) the interval adapted at every build, else 20.
int comm_interval(const mc_ctx *c) {
    if (c->rebuild_every > 0) return c->rebuild_every;
    return (c->comm && c->comm->peer_ok && c->halo_fused) ? c->comm->interval : 20;
}

extern "C" int mc_comm_halo_mode(mc_ctx *c, int *fused, char *why, int why_cap) {
    if (!c) return MC_E_INVALID;
    const bool on = c->comm && c->comm->peer_ok && c->halo_fused;
    if (fused) *fused = on ? 1 : 0;
    if (why && why_cap > 0) {
        const std::string w = !c->comm ? "no communicator" : (on ? "" : (c->halo_fused ? c->comm->peer_why : "option halo_fused = 0"));
        snprintf(why, (size_t)why_cap, "%s", w.c_str());
    }
    return MC_OK;
}

void comm_step_descriptors(mc_ctx *c, bool rebuild_step, HaloPush *push, HaloSplit *split) {
    CommState *cs = c->comm;
    const uint32_t e = ++cs->epoch;
    uint32_t *fl = cs->flags.p;
    HaloPush hp{};
    hp.n_first = (int)(cs->o_first_end - cs->o_own);
    hp.last_begin = (int)(cs->o_last_begin - cs->o_own);
    hp.to_prev = rebuild_step ? nullptr : cs->to_prev;
    hp.to_next = rebuild_step ? nullptr : cs->to_next;
    hp.ack_prev = fl + MC_HALO_ACK_FROM_PREV;
    hp.ack_next = fl + MC_HALO_ACK_FROM_NEXT;
    // towards prev this rank is "next", towards next it is "prev"
    hp.sig_ack_prev = cs->prev_map.flags + MC_HALO_ACK_FROM_NEXT;
    hp.sig_ack_next = cs->next_map.flags + MC_HALO_ACK_FROM_PREV;
    hp.sig_ready_prev = cs->prev_map.flags + MC_HALO_READY_FROM_NEXT;
    hp.sig_ready_next = cs->next_map.flags + MC_HALO_READY_FROM_PREV;
    hp.done_counter = fl + MC_HALO_CNT_KICK;
    hp.epoch = e;
    hp.err = c->rebuild_flag.p + 1;
    hp.max_disp2 = c->rebuild_flag.p + 3;
    *push = hp;
    split->n_first = hp.n_first;
    split->last_begin = hp.last_begin;
    split->wait.ready_prev = fl + MC_HALO_READY_FROM_PREV;
    split->wait.ready_next = fl + MC_HALO_READY_FROM_NEXT;
    split->wait.want = e;
    split->wait.err = c->rebuild_flag.p + 1;
}

int comm_rebuild(mc_ctx *c) {
    CommState *cs = c->comm;
    cudaStream_t st = c->st;
    struct Lap {  // MC_TRACE_STEP: wall clock of a rebuild that takes more than 5 ms, by phase (stall hunting)
        bool on; int dev; std::chrono::steady_clock::time_point t0, t; double ph[4] = {0, 0, 0, 0};
        void lap(int k) { if (!on) return; const auto now = std::chrono::steady_clock::now(); ph[k] += std::chrono::duration<double>(now - t).count(); t = now; }
        ~Lap() {
            if (!on) return;
            const double tot = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (tot > 5e-3) fprintf(stderr, "[comm_rebuild stall, device %d] %.2f ms: exchange launch %.2f, table all-gather + sync %.2f, rows %.2f, rest %.2f\n",
                                    dev, tot * 1e3, ph[0] * 1e3, ph[1] * 1e3, ph[2] * 1e3, ph[3] * 1e3);
        }
    } lapc{c->trace_step, c->device, std::chrono::steady_clock::now(), std::chrono::steady_clock::now()};
    if (c->grid_dirty) { int rc = dd_setup_grid(c); if (rc != MC_OK) return rc; }
    if (c->halo_fused && (!cs->peer_tried || (cs->peer_ok && (cs->exported[0] != c->xyzq[0].p || cs->exported[1] != c->xyzq[1].p)))) {
        int rc = peer_setup(c);
        if (rc != MC_OK) return rc;
    }
    TimedRegion tr(c, c->build_acc);
    const size_t cap = cs->cap;
    if ((size_t)c->n_rows > cap) { c->err = "domain decomposition: a rank owns more atoms than its exchange block holds"; return MC_E_CAPACITY; }
    const int prev = (cs->rank + cs->n - 1) % cs->n, next = (cs->rank + 1) % cs->n;
    MC_CUDAC(c, cs->d_layer.ensure(MC_DD_TABLE_WORDS));
    const size_t r0 = (size_t)c->row0;
    size_t n_all = 0;
    const int steps_in_interval = c->steps_since_build;
    if (cs->dd_migrate && cs->have_table) {
        // Neighbour-only migration.  Between two builds an atom moves less than a cell layer, so everything this
        // rank will hold afterwards -- owned layers and one ghost layer each side -- is owned now by itself, or
        // sits in the LAST TWO owned layers of the previous rank or the FIRST TWO of the next.  Those blocks are
        // contiguous in the cell-ordered arrays; their sizes are known on both sides from the table that the
        // previous build all-gathered, so there is no count exchange: own atoms are packed to the front of the
        // candidate arrays, the two blocks go out of them and the neighbours' blocks land right behind them.
        const uint32_t *me = cs->h_layer_all + MC_DD_TABLE_WORDS * (size_t)cs->rank;
        const uint32_t *tp = cs->h_layer_all + MC_DD_TABLE_WORDS * (size_t)prev, *tn = cs->h_layer_all + MC_DD_TABLE_WORDS * (size_t)next;
        const size_t n_own = (size_t)c->n_rows;
        // block ranges relative to the first owned slot; with two ranks and thin slabs the two blocks of a rank
        // overlap and go to the same peer: it gets the whole rank once
        auto blocks = [&](const uint32_t *t, size_t *lo_cnt, size_t *hi_off, size_t *hi_cnt) {
            const size_t own = t[4] - t[1];
            const bool whole = cs->n == 2 && t[8] > t[9];
            *lo_cnt = whole ? own : t[8] - t[1];   // to its prev
            *hi_off = whole ? own : t[9] - t[1];   // to its next
            *hi_cnt = whole ? 0 : t[4] - t[9];
        };
        size_t my_lo, my_hi_off, my_hi, p_lo, p_hi_off, p_hi, n_lo, n_hi_off, n_hi;
        blocks(me, &my_lo, &my_hi_off, &my_hi);
        blocks(tp, &p_lo, &p_hi_off, &p_hi);
        blocks(tn, &n_lo, &n_hi_off, &n_hi);
        // from next comes ITS low block (sent to its prev = me); from prev its high block.  With two ranks the
        // peer's "whole" arrives as its low block and its high block is empty.
        const size_t from_next = n_lo, from_prev = p_hi;
        n_all = n_own + from_next + from_prev;
        // Candidate order = [from prev | own | from next] (two ranks exchanging whole slabs: rank order).  Every
        // rank's array is then a sub-sequence of the same global sequence -- all ranks' owned blocks in ring
        // order -- so the stable sort leaves the atoms of a cell in the same relative order on the rank that
        // owns the cell and on the rank that holds it as a ghost layer: the per-step halo stays a plain block.
        const bool whole = cs->n == 2 && my_hi == 0 && from_prev == 0;
        const size_t off_own = whole ? (cs->rank == 0 ? 0 : from_next) : from_prev;
        const size_t off_next = whole ? (cs->rank == 0 ? n_own : 0) : from_prev + n_own;
        const size_t off_prev = 0;
        MC_CUDAC(c, cs->g_xyzq.ensure(n_all)); MC_CUDAC(c, cs->g_vel.ensure(n_all)); MC_CUDAC(c, cs->g_meta.ensure(n_all));
        MC_LAUNCH(dd_pack_kernel, div_up(n_own, 256), 256, 0, st, (int)n_own, (int)n_own, c->xyzq[c->cur].p + r0, c->vel[c->cur].p + r0,
                                                          c->orig[c->cur].p + r0, c->type[c->cur].p + r0, c->flags[c->cur].p + r0,
                                                          cs->g_xyzq.p + off_own, cs->g_vel.p + off_own, cs->g_meta.p + off_own);
        c->launches += 1;
        // NCCL pairs the k-th send to a peer with that peer's k-th receive from us: sends go (to prev, to next),
        // receives (from next, from prev), which also matches when prev == next
        MC_NCCL(c, nccl_api().GroupStart());
        if (my_lo) {
            MC_NCCL(c, nccl_api().Send(cs->g_xyzq.p + off_own, my_lo * 4, ncclFloat, prev, cs->comm, st));
            MC_NCCL(c, nccl_api().Send(cs->g_vel.p + off_own, my_lo * 4, ncclFloat, prev, cs->comm, st));
            MC_NCCL(c, nccl_api().Send(cs->g_meta.p + off_own, my_lo * 2, ncclInt32, prev, cs->comm, st));
        }
        if (my_hi) {
            MC_NCCL(c, nccl_api().Send(cs->g_xyzq.p + off_own + my_hi_off, my_hi * 4, ncclFloat, next, cs->comm, st));
            MC_NCCL(c, nccl_api().Send(cs->g_vel.p + off_own + my_hi_off, my_hi * 4, ncclFloat, next, cs->comm, st));
            MC_NCCL(c, nccl_api().Send(cs->g_meta.p + off_own + my_hi_off, my_hi * 2, ncclInt32, next, cs->comm, st));
        }
        if (from_next) {
            MC_NCCL(c, nccl_api().Recv(cs->g_xyzq.p + off_next, from_next * 4, ncclFloat, next, cs->comm, st));
            MC_NCCL(c, nccl_api().Recv(cs->g_vel.p + off_next, from_next * 4, ncclFloat, next, cs->comm, st));
            MC_NCCL(c, nccl_api().Recv(cs->g_meta.p + off_next, from_next * 2, ncclInt32, next, cs->comm, st));
        }
        if (from_prev) {
            MC_NCCL(c, nccl_api().Recv(cs->g_xyzq.p + off_prev, from_prev * 4, ncclFloat, prev, cs->comm, st));
            MC_NCCL(c, nccl_api().Recv(cs->g_vel.p + off_prev, from_prev * 4, ncclFloat, prev, cs->comm, st));
            MC_NCCL(c, nccl_api().Recv(cs->g_meta.p + off_prev, from_prev * 2, ncclInt32, prev, cs->comm, st));
        }
        MC_NCCL(c, nccl_api().GroupEnd());
    } else {
        // First build (atoms arrive as index blocks of the global system) and option dd_migrate = 0: every rank
        // contributes its atoms to an all-gather of fixed-capacity blocks and keeps what falls into its layers.
        n_all = cap * (size_t)cs->n;
        MC_CUDAC(c, cs->s_xyzq.ensure(cap)); MC_CUDAC(c, cs->s_vel.ensure(cap)); MC_CUDAC(c, cs->s_meta.ensure(cap));
        MC_CUDAC(c, cs->g_xyzq.ensure(n_all)); MC_CUDAC(c, cs->g_vel.ensure(n_all)); MC_CUDAC(c, cs->g_meta.ensure(n_all));
        MC_LAUNCH(dd_pack_kernel, div_up(cap, 256), 256, 0, st, (int)c->n_rows, (int)cap, c->xyzq[c->cur].p + r0, c->vel[c->cur].p + r0,
                                                         c->orig[c->cur].p + r0, c->type[c->cur].p + r0, c->flags[c->cur].p + r0,
                                                         cs->s_xyzq.p, cs->s_vel.p, cs->s_meta.p);
        c->launches += 1;
        MC_NCCL(c, nccl_api().GroupStart());
        MC_NCCL(c, nccl_api().AllGather(cs->s_xyzq.p, cs->g_xyzq.p, cap * 4, ncclFloat, cs->comm, st));
        MC_NCCL(c, nccl_api().AllGather(cs->s_vel.p, cs->g_vel.p, cap * 4, ncclFloat, cs->comm, st));
        MC_NCCL(c, nccl_api().AllGather(cs->s_meta.p, cs->g_meta.p, cap * 2, ncclInt32, cs->comm, st));
        MC_NCCL(c, nccl_api().GroupEnd());
        if (!cs->preconnected && cs->dd_migrate) {
            // NCCL connects its point-to-point channels lazily, at the first ncclSend / ncclRecv between two ranks --
            // tens of milliseconds that would otherwise land in the first neighbour-only migration, i.e. in the middle of
            // a run (round 1: a 27 ms stall ten steps in).  Pay for it here, with the message shape of a migration
            // (three sends + three receives per neighbour, ~1 MB each so that every channel a real block uses is up).
            const size_t m = 1u << 16;
            MC_CUDAC(c, cs->pre_buf.ensure(4 * m));
            MC_CUDAC(c, cudaMemsetAsync(cs->pre_buf.p, 0, sizeof(float4) * 2 * m, st));
            float4 *sb[2] = {cs->pre_buf.p, cs->pre_buf.p + m}, *rb[2] = {cs->pre_buf.p + 2 * m, cs->pre_buf.p + 3 * m};
            MC_NCCL(c, nccl_api().GroupStart());
            for (int k = 0; k < 3; ++k) MC_NCCL(c, nccl_api().Send(sb[0], m * 4, ncclFloat, prev, cs->comm, st));
            for (int k = 0; k < 3; ++k) MC_NCCL(c, nccl_api().Send(sb[1], m * 4, ncclFloat, next, cs->comm, st));
            for (int k = 0; k < 3; ++k) MC_NCCL(c, nccl_api().Recv(rb[0], m * 4, ncclFloat, next, cs->comm, st));
            for (int k = 0; k < 3; ++k) MC_NCCL(c, nccl_api().Recv(rb[1], m * 4, ncclFloat, prev, cs->comm, st));
            MC_NCCL(c, nccl_api().GroupEnd());
            cs->preconnected = true;
        }
    }
    cs->have_table = false;
    MC_CUDAC(c, c->keys[0].ensure(n_all)); MC_CUDAC(c, c->keys[1].ensure(n_all));
    MC_CUDAC(c, c->vals[0].ensure(n_all)); MC_CUDAC(c, c->vals[1].ensure(n_all));
    MC_CUDAC(c, c->scratch.ensure(std::max(radix_scratch_elems(n_all), scan_scratch_elems(n_all + 1)) + 64));
    MC_LAUNCH(dd_key_kernel, div_up(n_all, 256), 256, 0, st, (int)n_all, cs->g_xyzq.p, cs->g_meta.p, c->grid.p, c->keys[0].p, c->vals[0].p);
    c->launches += 1;
    uint32_t *kk[2] = {c->keys[0].p, c->keys[1].p}, *vv[2] = {c->vals[0].p, c->vals[1].p};
    const int which = radix_sort_pairs(kk, vv, n_all, c->key_bits, c->scratch.p, st, &c->launches);
    MC_CUDAC(c, cudaMemsetAsync(c->slot_of_orig.p, 0xff, sizeof(int) * (size_t)c->n_global, st));
    const int nx = c->cur ^ 1;
    const float r_list = std::max(c->rc_lj, c->rc_q) + c->skin;
    MC_LAUNCH(dd_reorder_kernel, div_up(n_all + 1, 256), 256, 0, st, 
        (int)n_all, kk[which], vv[which], c->grid.p, cs->g_xyzq.p, cs->g_vel.p, cs->g_meta.p, c->skin < 0.5f * r_list ? 1 : 0,
        c->xyzq[nx].p, c->xref.p, c->vel[nx].p, c->type[nx].p, c->flags[nx].p, c->orig[nx].p, c->slot_of_orig.p,
        c->cell_start.p, (uint32_t)cs->local_cap);
    const int plane = c->h_grid.nc[0] * c->h_grid.nc[1];
    MC_LAUNCH(dd_layer_offsets_kernel, 1, 32, 0, st, c->cell_start.p, plane, c->h_grid.nc[2], (uint32_t)nx, c->rebuild_flag.p + 3, cs->d_layer.p);
    c->launches += 2;
    // every rank's record travels with this rank's: the fused halo and the next migration need the neighbours'
    // block offsets (one 64-byte all-gather per rebuild)
    const size_t tw = MC_DD_TABLE_WORDS;
    MC_CUDAC(c, cs->d_layer_all.ensure(tw * (size_t)cs->n));
    if (!cs->h_layer_all) MC_CUDAC(c, cudaMallocHost(&cs->h_layer_all, sizeof(uint32_t) * tw * (size_t)cs->n));
    lapc.lap(0);
    MC_NCCL(c, nccl_api().AllGather(cs->d_layer.p, cs->d_layer_all.p, tw, ncclUint32, cs->comm, st));
    MC_CUDAC(c, cudaMemcpyAsync(cs->h_layer_all, cs->d_layer_all.p, sizeof(uint32_t) * tw * (size_t)cs->n, cudaMemcpyDeviceToHost, st));
    MC_CUDAC(c, cudaStreamSynchronize(st));
    lapc.lap(1);
    const uint32_t *h6 = cs->h_layer_all + tw * (size_t)cs->rank;
    cs->o_gp = h6[0]; cs->o_own = h6[1]; cs->o_first_end = h6[2]; cs->o_last_begin = h6[3]; cs->o_own_end = h6[4]; cs->o_end = h6[5];
    if (cs->o_end > cs->local_cap) { c->err = "domain decomposition: local atom capacity exceeded"; return MC_E_CAPACITY; }
    {
        // every atom must have exactly one owner after the exchange; the same table gives the largest displacement
        // of the interval that just ended, which sets the next interval when the schedule is adaptive
        int64_t owned = 0;
        float d2 = 0.f;
        for (int r = 0; r < cs->n; ++r) {
            const uint32_t *t = cs->h_layer_all + tw * (size_t)r;
            owned += (int64_t)t[4] - (int64_t)t[1];
            float v;
            memcpy(&v, &t[7], sizeof(float));
            d2 = std::max(d2, v);
        }
        if (owned != c->n_global) {
            c->err = "domain decomposition: " + std::to_string(owned) + " atoms have an owner after the exchange, the system has " +
                     std::to_string(c->n_global) + " (an atom crossed more than one cell layer between two builds)";
            return MC_E_COMM;
        }
        const double half_skin = 0.5 * (double)c->skin;
        if (steps_in_interval > 0 && half_skin > 0.0) {
            const double frac = std::sqrt((double)d2) / half_skin;
            cs->last_disp_frac = frac;
            // the fastest atom moves ballistically over one interval: aim at 75 % of skin/2 (the largest displacement of
            // an interval fluctuates by 10-20 % from one interval to the next, more on small systems: replayed on oracle
            // trajectories in tests/test_rebuild_flag_model.py, 85 % overshoots now and then, 75 % does not), at most double
            // (the extrapolation is linear in time, an upper bound for anything slower than ballistic motion: doubling is safe)
            double want = frac > 1e-6 ? 0.75 * steps_in_interval / frac : 2.0 * steps_in_interval + 1;
            want = std::min(want, 2.0 * steps_in_interval + 1.0);
            cs->interval = (int)std::max(4.0, std::min(200.0, std::floor(want)));
            // every rank reads the same table: an atom that outran skin/2 anywhere is counted everywhere
            if (frac > 1.0) c->n_list_violations++;
        }
    }
    cs->have_table = true;
    if (cs->peer_ok) {
        const uint32_t *hp = cs->h_layer_all + tw * (size_t)prev;
        const uint32_t *hn = cs->h_layer_all + tw * (size_t)next;
        // my first owned layer is prev's ghost layer behind its owned slots; my last one is next's ghost layer in front
        if (hp[5] - hp[4] != cs->o_first_end - cs->o_own || hn[1] - hn[0] != cs->o_own_end - cs->o_last_begin) {
            c->err = "domain decomposition: a neighbour's ghost block and this rank's boundary layer differ in size";
            return MC_E_COMM;
        }
        cs->to_prev = cs->prev_map.xyzq[hp[6] & 1] + hp[4];
        cs->to_next = cs->next_map.xyzq[hn[6] & 1] + hn[0];
    }
    c->cur = nx;
    c->cell_of_slot = kk[which];
    c->n = cs->o_end;
    c->row0 = cs->o_own;
    c->n_rows = cs->o_own_end - cs->o_own;
    c->identity_order = false;
    tr.stop();
    lapc.lap(3);
    const int rc_rows = engine_build_rows(c);
    lapc.lap(2);
    return rc_rows;
}

// Per-step ghost refresh: boundary blocks out, ghost blocks in, straight from / into xyzq.
int comm_halo_positions(mc_ctx *c) {
    CommState *cs = c->comm;
    TimedRegion tr(c, c->halo_acc);
    float4 *x = c->xyzq[c->cur].p;
    const int prev = (cs->rank + cs->n - 1) % cs->n, next = (cs->rank + 1) % cs->n;
    const size_t n_first = cs->o_first_end - cs->o_own, n_last = cs->o_own_end - cs->o_last_begin;
    const size_t n_gp = cs->o_own - cs->o_gp, n_gn = cs->o_end - cs->o_own_end;
    // with two ranks prev == next: NCCL pairs the k-th send to a peer with that peer's k-th recv from us,
    // so the receive order (from next, then from prev) mirrors the send order (to prev, then to next)
    MC_NCCL(c, nccl_api().GroupStart());
    MC_NCCL(c, nccl_api().Send(x + cs->o_own, n_first * 4, ncclFloat, prev, cs->comm, c->st));
    MC_NCCL(c, nccl_api().Send(x + cs->o_last_begin, n_last * 4, ncclFloat, next, cs->comm, c->st));
    MC_NCCL(c, nccl_api().Recv(x + cs->o_own_end, n_gn * 4, ncclFloat, next, cs->comm, c->st));
    MC_NCCL(c, nccl_api().Recv(x + cs->o_gp, n_gp * 4, ncclFloat, prev, cs->comm, c->st));
    MC_NCCL(c, nccl_api().GroupEnd());
    tr.stop();
    return MC_OK;
}

int comm_agree_flag(mc_ctx *c, bool *flag) {
    CommState *cs = c->comm;
    MC_CUDAC(c, cs->d_red.ensure(4));
    int *d = reinterpret_cast<int *>(cs->d_red.p);
    const int v = *flag ? 1 : 0;
    MC_CUDAC(c, cudaMemcpyAsync(d, &v, sizeof(int), cudaMemcpyHostToDevice, c->st));
    MC_NCCL(c, nccl_api().AllReduce(d, d, 1, ncclInt32, ncclMax, cs->comm, c->st));
    int out = 0;
    MC_CUDAC(c, cudaMemcpyAsync(&out, d, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    MC_CUDAC(c, cudaStreamSynchronize(c->st));
    *flag = out != 0;
    return MC_OK;
}

// max over ranks of the two flag words behind the step kernels {displacement / non-finite bits, halo error bits},
// delivered to pinned host memory; no synchronisation here -- the caller's one cudaStreamSynchronize covers it
int comm_reduce_flags_async(mc_ctx *c, const int *d_flags2, int *h_out2) {
    CommState *cs = c->comm;
    MC_CUDAC(c, cs->d_red.ensure(4));
    int *d = reinterpret_cast<int *>(cs->d_red.p);
    MC_CUDAC(c, cudaMemcpyAsync(d, d_flags2, 2 * sizeof(int), cudaMemcpyDeviceToDevice, c->st));
    MC_NCCL(c, nccl_api().AllReduce(d, d, 2, ncclInt32, ncclMax, cs->comm, c->st));
    MC_CUDAC(c, cudaMemcpyAsync(h_out2, d, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    return MC_OK;
}

void comm_shrink_interval(mc_ctx *c) {
    if (c->comm) c->comm->interval = std::max(4, c->comm->interval / 2);
}

int comm_allreduce3(mc_ctx *c, double v[3]) {
    CommState *cs = c->comm;
    MC_CUDAC(c, cs->d_red.ensure(4));
    MC_CUDAC(c, cudaMemcpyAsync(cs->d_red.p, v, 3 * sizeof(double), cudaMemcpyHostToDevice, c->st));
    MC_NCCL(c, nccl_api().AllReduce(cs->d_red.p, cs->d_red.p, 3, ncclFloat64, ncclSum, cs->comm, c->st));
    MC_CUDAC(c, cudaMemcpyAsync(v, cs->d_red.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    MC_CUDAC(c, cudaStreamSynchronize(c->st));
    return MC_OK;
}

// in-place sum over the ranks of n doubles in device memory, on the engine stream (no host synchronisation): the kinetic
// energy the CSVR thermostat needs inside the step
int comm_allreduce_dev_f64(mc_ctx *c, double *d, int n) {
    CommState *cs = c->comm;
    MC_NCCL(c, nccl_api().AllReduce(d, d, (size_t)n, ncclFloat64, ncclSum, cs->comm, c->st));
    return MC_OK;
}

void comm_rank_size(const mc_ctx *c, int *rank, int *n_ranks) {
    *rank = c->comm ? c->comm->rank : 0;
    *n_ranks = c->comm ? c->comm->n : 1;
}

// The external-force blocks of all ranks, and -- in the same NCCL group, i.e. the same launch -- every rank's two flag words
// {displacement / non-finite bits, halo error bits} as they stand on the stream (those of the previous step's kick_drift).
// The per-rank words land in pinned host memory h_flags[2 * n_ranks] behind the call's one synchronisation.
int comm_allgather_ext_and_flags(mc_ctx *c, float *buf, size_t chunk, const int *d_flags2, int *h_flags) {
    CommState *cs = c->comm;
    MC_CUDAC(c, cs->d_flags_all.ensure(2 * (size_t)cs->n + 2));
    int *mine = cs->d_flags_all.p + 2 * (size_t)cs->n;
    MC_CUDAC(c, cudaMemcpyAsync(mine, d_flags2, 2 * sizeof(int), cudaMemcpyDeviceToDevice, c->st));
    MC_NCCL(c, nccl_api().GroupStart());
    MC_NCCL(c, nccl_api().AllGather(buf + (size_t)cs->rank * chunk, buf, chunk, ncclFloat, cs->comm, c->st));
    MC_NCCL(c, nccl_api().AllGather(mine, cs->d_flags_all.p, 2, ncclInt32, cs->comm, c->st));
    MC_NCCL(c, nccl_api().GroupEnd());
    MC_CUDAC(c, cudaMemcpyAsync(h_flags, cs->d_flags_all.p, 2 * sizeof(int) * (size_t)cs->n, cudaMemcpyDeviceToHost, c->st));
    return MC_OK;
}

int comm_allgather_f32_inplace(mc_ctx *c, float *buf, size_t chunk) {
    CommState *cs = c->comm;
    MC_NCCL(c, nccl_api().AllGather(buf + (size_t)cs->rank * chunk, buf, chunk, ncclFloat, cs->comm, c->st));
    return MC_OK;
}

int comm_allreduce_f4(mc_ctx *c, float4 *buf, int64_t n) {
    CommState *cs = c->comm;
    MC_NCCL(c, nccl_api().AllReduce(buf, buf, (size_t)n * 4, ncclFloat, ncclSum, cs->comm, c->st));
    return MC_OK;
}
