// comm.cu -- spatial domain decomposition over the GPUs of one box: slabs of whole cell
// layers along z, one ghost layer on each side, NCCL over NVLink/NVSwitch.
//
// Layout on a rank (cell order, z slowest):   [ ghost layer from prev | owned layers | ghost layer from next ]
// Because the cell sort is z-major, each of the three parts -- and the first / last OWNED layer that
// the neighbours need as their ghosts -- is one contiguous block of the position array.  The
// per-step halo exchange therefore needs no pack / unpack kernels: each rank ncclSend's the two
// boundary blocks straight out of xyzq and ncclRecv's the two ghost blocks straight into xyzq
// (0.4-0.9 MB per rank per step on the 1M-atom fluid).  Full (not half) neighbour lists mean every
// rank computes the forces of its own atoms only: there is no reverse (force) communication.
//
// Positions are kept in the GLOBAL frame and the minimum image is applied to pair differences,
// exactly as on one GPU (ghosts are never shifted by +-L, which would cost ~1e-5 A of fp32
// resolution at |z| ~ L), so a decomposed run lists and masks the same pairs as the single-GPU run.
//
// Rebuild (every rebuild_every steps): every rank contributes its owned atoms to an ncclAllGather of
// fixed-capacity blocks (padding marked with id -1), then keeps what falls into its layer range
// [kz0-1, kz1] by a filter key + the same radix sort / reorder as the single-GPU path.  No count
// exchange, no host round trip, no migration bookkeeping; 40 MB through NVSwitch every ~20 steps
// on the 1M-atom system.  (A neighbour-only migration is the obvious next refinement for systems
// far beyond 1e7 atoms.)
#include "nccl_dyn.cuh"

#include <algorithm>
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
    // slot offsets of the current build
    uint32_t o_gp = 0, o_own = 0, o_first_end = 0, o_last_begin = 0, o_own_end = 0, o_end = 0;
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
    if (g_meta[i].x < 0) { keys[i] = (uint32_t)g.ncell; return; }
    float4 p = g_xyzq[i];
    float c[3] = {p.x, p.y, p.z};
    int k[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        c[a] -= floorf((c[a] - g.lo[a]) * g.inv_ext[a]) * g.ext[a];
        if (c[a] < g.lo[a]) c[a] += g.ext[a];
        if (c[a] >= g.lo[a] + g.ext[a]) c[a] -= g.ext[a];
        const int nca = a == 2 ? g.ncz_global : g.nc[a];
        const int kk = (int)floorf((c[a] - g.lo[a]) * g.inv_cw[a]);
        k[a] = min(max(kk, 0), nca - 1);
    }
    g_xyzq[i] = make_float4(c[0], c[1], c[2], p.w);
    const int l = (k[2] - g.kz_off + g.ncz_global) % g.ncz_global;  // local layer of this global layer
    keys[i] = l < g.nc[2] ? (uint32_t)((l * g.nc[1] + k[1]) * g.nc[0] + k[0]) : (uint32_t)g.ncell;
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
    const int ncell = gp->ncell;
    const int prev = k == 0 ? -1 : (int)skeys[k - 1];
    const int cur = k == n_all ? ncell + 1 : (int)skeys[k];
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

__global__ void dd_layer_offsets_kernel(const uint32_t *__restrict__ cell_start, int plane, int nlayers,
                                        uint32_t *__restrict__ out6) {
    if (threadIdx.x != 0) return;
    out6[0] = cell_start[0];
    out6[1] = cell_start[plane];                  // first owned slot
    out6[2] = cell_start[2 * plane];              // end of the first owned layer
    out6[3] = cell_start[(nlayers - 2) * plane];  // begin of the last owned layer
    out6[4] = cell_start[(nlayers - 1) * plane];  // end of the owned slots
    out6[5] = cell_start[nlayers * plane];        // end of the local slots
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

void comm_destroy(mc_ctx *c) {
    if (!c->comm) return;
    CommState *cs = c->comm;
    if (cs->comm) nccl_api().CommDestroy(cs->comm);
    cs->s_xyzq.release(); cs->s_vel.release(); cs->g_xyzq.release(); cs->g_vel.release();
    cs->s_meta.release(); cs->g_meta.release(); cs->d_layer.release(); cs->d_red.release();
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
    MC_CUDAC(c, cudaMemset(c->slot_of_orig.p, 0xff, sizeof(int) * (size_t)n));
    return MC_OK;
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
    c->h_grid = g;
    c->ncell_cap = (size_t)g.ncell;
    MC_CUDAC(c, c->grid.ensure(1));
    MC_CUDAC(c, cudaMemcpyAsync(c->grid.p, &c->h_grid, sizeof(GridParams), cudaMemcpyHostToDevice, c->st));
    MC_CUDAC(c, c->cell_start.ensure(c->ncell_cap + 2));
    int bits = 1;
    while (((size_t)1 << bits) < c->ncell_cap + 1) ++bits;  // + sentinel key
    c->key_bits = bits;
    c->grid_dirty = false;
    return MC_OK;
}

int comm_rebuild(mc_ctx *c) {
    CommState *cs = c->comm;
    cudaStream_t st = c->st;
    if (c->grid_dirty) { int rc = dd_setup_grid(c); if (rc != MC_OK) return rc; }
    TimedRegion tr(c, c->build_acc);
    const size_t cap = cs->cap, n_all = cap * (size_t)cs->n;
    if ((size_t)c->n_rows > cap) { c->err = "domain decomposition: a rank owns more atoms than its all-gather block holds"; return MC_E_CAPACITY; }
    MC_CUDAC(c, cs->s_xyzq.ensure(cap)); MC_CUDAC(c, cs->s_vel.ensure(cap)); MC_CUDAC(c, cs->s_meta.ensure(cap));
    MC_CUDAC(c, cs->g_xyzq.ensure(n_all)); MC_CUDAC(c, cs->g_vel.ensure(n_all)); MC_CUDAC(c, cs->g_meta.ensure(n_all));
    MC_CUDAC(c, cs->d_layer.ensure(8));
    const size_t r0 = (size_t)c->row0;
    dd_pack_kernel<<<div_up(cap, 256), 256, 0, st>>>((int)c->n_rows, (int)cap, c->xyzq[c->cur].p + r0, c->vel[c->cur].p + r0,
                                                     c->orig[c->cur].p + r0, c->type[c->cur].p + r0, c->flags[c->cur].p + r0,
                                                     cs->s_xyzq.p, cs->s_vel.p, cs->s_meta.p);
    c->launches += 1;
    MC_NCCL(c, nccl_api().GroupStart());
    MC_NCCL(c, nccl_api().AllGather(cs->s_xyzq.p, cs->g_xyzq.p, cap * 4, ncclFloat, cs->comm, st));
    MC_NCCL(c, nccl_api().AllGather(cs->s_vel.p, cs->g_vel.p, cap * 4, ncclFloat, cs->comm, st));
    MC_NCCL(c, nccl_api().AllGather(cs->s_meta.p, cs->g_meta.p, cap * 2, ncclInt32, cs->comm, st));
    MC_NCCL(c, nccl_api().GroupEnd());
    MC_CUDAC(c, c->keys[0].ensure(n_all)); MC_CUDAC(c, c->keys[1].ensure(n_all));
    MC_CUDAC(c, c->vals[0].ensure(n_all)); MC_CUDAC(c, c->vals[1].ensure(n_all));
    MC_CUDAC(c, c->scratch.ensure(std::max(radix_scratch_elems(n_all), scan_scratch_elems(n_all + 1)) + 64));
    dd_key_kernel<<<div_up(n_all, 256), 256, 0, st>>>((int)n_all, cs->g_xyzq.p, cs->g_meta.p, c->grid.p, c->keys[0].p, c->vals[0].p);
    c->launches += 1;
    uint32_t *kk[2] = {c->keys[0].p, c->keys[1].p}, *vv[2] = {c->vals[0].p, c->vals[1].p};
    const int which = radix_sort_pairs(kk, vv, n_all, c->key_bits, c->scratch.p, st, &c->launches);
    MC_CUDAC(c, cudaMemsetAsync(c->slot_of_orig.p, 0xff, sizeof(int) * (size_t)c->n_global, st));
    const int nx = c->cur ^ 1;
    const float r_list = std::max(c->rc_lj, c->rc_q) + c->skin;
    dd_reorder_kernel<<<div_up(n_all + 1, 256), 256, 0, st>>>(
        (int)n_all, kk[which], vv[which], c->grid.p, cs->g_xyzq.p, cs->g_vel.p, cs->g_meta.p, c->skin < 0.5f * r_list ? 1 : 0,
        c->xyzq[nx].p, c->xref.p, c->vel[nx].p, c->type[nx].p, c->flags[nx].p, c->orig[nx].p, c->slot_of_orig.p,
        c->cell_start.p, (uint32_t)cs->local_cap);
    const int plane = c->h_grid.nc[0] * c->h_grid.nc[1];
    dd_layer_offsets_kernel<<<1, 32, 0, st>>>(c->cell_start.p, plane, c->h_grid.nc[2], cs->d_layer.p);
    c->launches += 2;
    uint32_t *h6 = reinterpret_cast<uint32_t *>(c->h_pinned) + 16;
    MC_CUDAC(c, cudaMemcpyAsync(h6, cs->d_layer.p, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    MC_CUDAC(c, cudaStreamSynchronize(st));
    cs->o_gp = h6[0]; cs->o_own = h6[1]; cs->o_first_end = h6[2]; cs->o_last_begin = h6[3]; cs->o_own_end = h6[4]; cs->o_end = h6[5];
    if (cs->o_end > cs->local_cap) { c->err = "domain decomposition: local atom capacity exceeded"; return MC_E_CAPACITY; }
    c->cur = nx;
    c->cell_of_slot = kk[which];
    c->n = cs->o_end;
    c->row0 = cs->o_own;
    c->n_rows = cs->o_own_end - cs->o_own;
    c->identity_order = false;
    tr.stop();
    return engine_build_rows(c);
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

int comm_allreduce3(mc_ctx *c, double v[3]) {
    CommState *cs = c->comm;
    MC_CUDAC(c, cs->d_red.ensure(4));
    MC_CUDAC(c, cudaMemcpyAsync(cs->d_red.p, v, 3 * sizeof(double), cudaMemcpyHostToDevice, c->st));
    MC_NCCL(c, nccl_api().AllReduce(cs->d_red.p, cs->d_red.p, 3, ncclFloat64, ncclSum, cs->comm, c->st));
    MC_CUDAC(c, cudaMemcpyAsync(v, cs->d_red.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    MC_CUDAC(c, cudaStreamSynchronize(c->st));
    return MC_OK;
}

int comm_allreduce_f4(mc_ctx *c, float4 *buf, int64_t n) {
    CommState *cs = c->comm;
    MC_NCCL(c, nccl_api().AllReduce(buf, buf, (size_t)n * 4, ncclFloat, ncclSum, cs->comm, c->st));
    return MC_OK;
}
