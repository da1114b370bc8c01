// engine.cu -- the C ABI of include/molchanica_md.h: handle lifetime, system upload, the
// rebuild / force / step orchestration on the handle's stream, and read-back.
//
// Internal data layout (all device-resident, "cell order" = the order produced by the last
// neighbour build, tracked by orig[] / slot_of_orig[]):
//   xyzq[2]   float4  x, y, z, q          (ping-pong across reorders)
//   vel[2]    float4  vx, vy, vz, 1/m
//   force     float4  fx, fy, fz, e_i
//   xref      float4  positions at the last build (displacement criterion)
//   type[2] u16, flags[2] u8, orig[2] i32, slot_of_orig i32
//   cell_start u32[ncell+1]; nbr_count u32[n]; nbr_start u32[n+1] (rows padded to 8 entries);
//   nbr_list u32[capacity]
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

#include "common.cuh"
#include "csvr_terms.h"
#include "dock.cuh"
#include "engine.cuh"
#include "md_fused.cuh"
#include "integrate.cuh"
#include "neighbor.cuh"
#include "pair_force.cuh"

static std::string g_create_err;

#define MC_CUDA(ctx, call)                                                                                 \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +          \
                         std::to_string(__LINE__) + ")";                                                   \
            return MC_E_CUDA;                                                                              \
        }                                                                                                  \
    } while (0)

#define MC_REQUIRE(ctx, cond, msg) \
    do {                           \
        if (!(cond)) {             \
            (ctx)->err = (msg);    \
            return MC_E_INVALID;   \
        }                          \
    } while (0)

static int fail(mc_ctx *c, int code, const std::string &m) {
    c->err = m;
    return code;
}

// see engine_flush_tail
#define MC_FLUSH_OBS(c)                               \
    do {                                              \
        if ((c)->tail_pending) {                      \
            const int rc_flush_ = engine_flush_tail(c); \
            if (rc_flush_ != MC_OK) return rc_flush_; \
        }                                             \
    } while (0)
// setters (anything that may move, reorder or redefine atoms) also drop the fused kernel's private list; pure observers
// (MC_FLUSH_OBS) leave it alone -- a list build they trigger reorders the atoms and drops it in engine_build_list
#define MC_FLUSH(c)                                   \
    do {                                              \
        (c)->bl_valid = false;                        \
        MC_FLUSH_OBS(c);                              \
    } while (0)

// ---- lifetime ------------------------------------------------------------------------------------

extern "C" int mc_abi_version(void) { return MC_ABI_VERSION; }

extern "C" int mc_struct_sizes(int *energy_bytes, int *stats_bytes) {
    if (energy_bytes) *energy_bytes = (int)sizeof(mc_energy);
    if (stats_bytes) *stats_bytes = (int)sizeof(mc_stats);
    return MC_OK;
}

extern "C" const char *mc_last_error(const mc_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" int mc_create(int device, mc_ctx **out) {
    if (!out) { g_create_err = "mc_create: out is NULL"; return MC_E_INVALID; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        g_create_err = std::string("mc_create: no CUDA device (") + cudaGetErrorString(e) +
                       "); this engine has no CPU fallback";
        return MC_E_NODEVICE;
    }
    if (device < 0 || device >= count) {
        g_create_err = "mc_create: device ordinal " + std::to_string(device) + " out of range (" + std::to_string(count) + " devices)";
        return MC_E_NODEVICE;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_err = std::string("mc_create: ") + cudaGetErrorString(e);
        return MC_E_NODEVICE;
    }
    if (prop.major != 10) {
        g_create_err = "mc_create: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                       "; this library carries sm_100a code only";
        return MC_E_NODEVICE;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_err = std::string("mc_create: cudaSetDevice: ") + cudaGetErrorString(e);
        return MC_E_NODEVICE;
    }
    mc_ctx *c = new mc_ctx();
    c->device = device;
    c->n_sms = prop.multiProcessorCount;
    c->l2_bytes = (size_t)prop.l2CacheSize;
    if ((e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = pair_force_prepare()) != cudaSuccess || (e = dock_prepare()) != cudaSuccess ||
        (e = tile_sweep_prepare()) != cudaSuccess || (e = pair_tile_prepare()) != cudaSuccess || (e = md_fused_prepare()) != cudaSuccess ||
        (e = cudaMallocHost(&c->h_pinned, 256)) != cudaSuccess) {
        g_create_err = std::string("mc_create: ") + cudaGetErrorString(e);
        delete c;
        return MC_E_CUDA;
    }
    c->trace_step = getenv("MC_TRACE_STEP") != nullptr;
    const char *ln = getenv("MC_PAIR_LANES");
    if (ln) c->pair_lanes = atoi(ln);
    *out = c;
    return MC_OK;
}

extern "C" int mc_destroy(mc_ctx *c) {
    if (!c) return MC_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->st);
    if (c->trace_step && c->trace_calls > 0) {
        static const char *nm[8] = {"upload+setup", "close_tail/forces", "wait+allgather", "step loop", "flags", "sync", "after", ""};
        fprintf(stderr, "[mc_step trace, device %d, %lld calls] ms per call:", c->device, (long long)c->trace_calls);
        for (int k = 0; k < 7; ++k) fprintf(stderr, " %s %.4f", nm[k], c->trace_t[k] / (double)c->trace_calls * 1e3);
        fprintf(stderr, " | device ms per call: upload %.4f tail(forces+half kick) %.4f all-gather %.4f kick+drift %.4f\n",
                c->trace_dev[0] / (double)c->trace_calls, c->trace_dev[1] / (double)c->trace_calls, c->trace_dev[2] / (double)c->trace_calls,
                c->trace_dev[3] / (double)c->trace_calls);
        for (int k = 0; k < 8; ++k) if (c->trace_ev[k]) cudaEventDestroy(c->trace_ev[k]);
    }
    comm_destroy(c);
    pme_release(&c->pme);
    for (auto &p : c->ev_pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    if (c->ev_step_a) {
        cudaEventDestroy(c->ev_step_a); cudaEventDestroy(c->ev_step_b);
        cudaEventDestroy(c->ev_flag[0]); cudaEventDestroy(c->ev_flag[1]);
    }
    if (c->st_up) { cudaStreamSynchronize(c->st_up); cudaEventDestroy(c->ev_up); cudaStreamDestroy(c->st_up); }
    if (c->ev_drift) cudaEventDestroy(c->ev_drift);
    if (c->st_copy) {
        cudaStreamSynchronize(c->st_copy);
        for (int b = 0; b < 2; ++b) { cudaEventDestroy(c->ev_snap_staged[b]); cudaEventDestroy(c->ev_snap_done[b]); }
        cudaStreamDestroy(c->st_copy);
    }
    c->free_all();
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->h_flags_all) cudaFreeHost(c->h_flags_all);
    cudaStreamDestroy(c->st);
    delete c;
    return MC_OK;
}

// ---- system definition ---------------------------------------------------------------------------

extern "C" int mc_set_box(mc_ctx *c, const float lo[3], const float hi[3], int periodic) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    MC_REQUIRE(c, lo && hi, "mc_set_box: NULL bounds");
    for (int a = 0; a < 3; ++a) {
        MC_REQUIRE(c, !periodic || hi[a] > lo[a], "mc_set_box: periodic box needs hi > lo on every axis");
        c->lo[a] = lo[a];
        c->ext[a] = hi[a] - lo[a];
    }
    c->periodic = periodic != 0;
    c->grid_dirty = true;
    c->list_valid = false;
    c->forces_valid = false;
    return MC_OK;
}

extern "C" int mc_set_cutoffs(mc_ctx *c, float rc_lj, float rc_q, float skin, int coulomb_mode, float alpha) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    MC_REQUIRE(c, rc_lj > 0.f && rc_q > 0.f && skin >= 0.f, "mc_set_cutoffs: cutoffs must be positive, skin >= 0");
    MC_REQUIRE(c, coulomb_mode >= MC_COULOMB_NONE && coulomb_mode <= MC_COULOMB_ERFC, "mc_set_cutoffs: bad coulomb_mode");
    c->rc_lj = rc_lj; c->rc_q = rc_q; c->skin = skin; c->coul_mode = coulomb_mode; c->alpha = alpha;
    c->grid_dirty = true;
    c->list_valid = false;
    c->forces_valid = false;
    return MC_OK;
}

extern "C" int mc_set_overrides(mc_ctx *c, int lj_disabled, int coulomb_disabled) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    c->lj_disabled = lj_disabled != 0;
    c->coul_disabled = coulomb_disabled != 0;
    c->forces_valid = false;
    return MC_OK;
}

extern "C" int mc_set_lj_table(mc_ctx *c, int n_types, const float *sigma_eps) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    MC_REQUIRE(c, n_types >= 1 && n_types <= pair_force_max_types() && sigma_eps,
               "mc_set_lj_table: 1 <= n_types <= " + std::to_string(pair_force_max_types()) + " and a table are required");
    cudaSetDevice(c->device);
    std::vector<float2> t((size_t)n_types * n_types);
    for (size_t k = 0; k < t.size(); ++k) {
        const float s = sigma_eps[2 * k], e = sigma_eps[2 * k + 1];
        t[k] = make_float2(s * s, 24.f * e);
    }
    MC_CUDA(c, c->ljtab.ensure(t.size()));
    MC_CUDA(c, cudaMemcpyAsync(c->ljtab.p, t.data(), t.size() * sizeof(float2), cudaMemcpyHostToDevice, c->st));
    MC_CUDA(c, cudaStreamSynchronize(c->st));
    c->n_types = n_types;
    c->sig2_0 = t[0].x;
    c->eps24_0 = t[0].y;
    c->forces_valid = false;
    return MC_OK;
}

static int upload_atoms_local(mc_ctx *c, int64_t n, const mc_float4 *xyzq, const uint16_t *type, const mc_float4 *vel,
                              const uint8_t *flags, const int *orig_ids, size_t alloc_n) {
    // allocate every per-atom array for alloc_n >= n local atoms and upload n in the given order
    MC_CUDA(c, c->alloc_atoms(std::max<size_t>((size_t)n, alloc_n)));
    c->n = n;
    c->cur = 0;
    cudaStream_t st = c->st;
    std::vector<uint16_t> ty;
    std::vector<uint8_t> fl;
    std::vector<mc_float4> v;
    std::vector<int> ids;
    if (!type) { ty.assign((size_t)n, 0); type = ty.data(); }
    if (!flags) { fl.assign((size_t)n, 0); flags = fl.data(); }
    if (!vel) { v.assign((size_t)n, mc_float4{0.f, 0.f, 0.f, 1.f}); vel = v.data(); }
    if (!orig_ids) { ids.resize((size_t)n); for (int64_t k = 0; k < n; ++k) ids[(size_t)k] = (int)k; orig_ids = ids.data(); }
    if (n > 0) {
        MC_CUDA(c, cudaMemcpyAsync(c->xyzq[0].p, xyzq, n * sizeof(float4), cudaMemcpyHostToDevice, st));
        MC_CUDA(c, cudaMemcpyAsync(c->vel[0].p, vel, n * sizeof(float4), cudaMemcpyHostToDevice, st));
        MC_CUDA(c, cudaMemcpyAsync(c->type[0].p, type, n * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
        MC_CUDA(c, cudaMemcpyAsync(c->flags[0].p, flags, n * sizeof(uint8_t), cudaMemcpyHostToDevice, st));
        MC_CUDA(c, cudaMemcpyAsync(c->orig[0].p, orig_ids, n * sizeof(int), cudaMemcpyHostToDevice, st));
        MC_CUDA(c, cudaMemsetAsync(c->force.p, 0, n * sizeof(float4), st));
    }
    MC_CUDA(c, cudaStreamSynchronize(st));
    c->identity_order = true;
    return MC_OK;
}

extern "C" int mc_set_atoms(mc_ctx *c, int64_t n, const mc_float4 *xyzq, const uint16_t *type,
                            const mc_float4 *vel_invmass, const uint8_t *flags) {
    if (!c) return MC_E_INVALID;
    MC_REQUIRE(c, n >= 0 && n < (int64_t)1 << 31, "mc_set_atoms: 0 <= n < 2^31 required");
    MC_REQUIRE(c, n == 0 || xyzq, "mc_set_atoms: xyzq is NULL");
    cudaSetDevice(c->device);
    c->n_global = n;
    c->bl_valid = false;
    c->tail_pending = false;  // a new system: nothing of the old one is left to finish
    c->cons_virial_valid = false;
    c->list_valid = false;
    c->forces_valid = false;
    c->have_excl = false;
    c->have_p14 = false;
    c->n_bonds = c->n_angles = c->n_dihedrals = 0;
    c->n_waters = 0;
    c->n_vsites = 0;
    c->n_hclusters = c->n_hconstraints = 0;
    c->h_in_water.clear();
    c->h_in_hcluster.clear();
    c->have_mols = false;
    c->n_pairs_listed = 0;
    c->total_mass = 0.0;
    c->pme.self_q2 = 0.0;
    for (int64_t k = 0; k < n; ++k) c->pme.self_q2 += (double)xyzq[k].w * (double)xyzq[k].w;
    for (int64_t k = 0; k < n; ++k) {
        const float im = vel_invmass ? vel_invmass[k].w : 1.f;
        if (im > 0.f) c->total_mass += 1.0 / (double)im;
    }
    if (c->comm_active) return comm_set_atoms(c, n, xyzq, type, vel_invmass, flags);
    c->n_rows = n;
    if (type)
        for (int64_t k = 0; k < n; ++k) MC_REQUIRE(c, type[k] < pair_force_max_types(), "mc_set_atoms: type id out of range");
    int rc = upload_atoms_local(c, n, xyzq, type, vel_invmass, flags, nullptr, (size_t)n);
    if (rc != MC_OK) return rc;
    // slot_of_orig = identity
    std::vector<int> ids((size_t)n);
    for (int64_t k = 0; k < n; ++k) ids[(size_t)k] = (int)k;
    if (n) MC_CUDA(c, cudaMemcpy(c->slot_of_orig.p, ids.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    return MC_OK;
}

int engine_upload_local(mc_ctx *c, int64_t n, const mc_float4 *xyzq, const uint16_t *type, const mc_float4 *vel,
                        const uint8_t *flags, const int *orig_ids, size_t alloc_n) {
    return upload_atoms_local(c, n, xyzq, type, vel, flags, orig_ids, alloc_n);
}

extern "C" int mc_set_exclusions(mc_ctx *c, const int32_t *start, const int32_t *idx) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    cudaSetDevice(c->device);
    c->list_valid = false;
    c->forces_valid = false;
    if (!start || !idx || start[c->n_global] == 0) { c->have_excl = false; return MC_OK; }
    const int64_t n = c->n_global, m = start[n];
    MC_REQUIRE(c, m >= 0, "mc_set_exclusions: negative total");
    // the rows are dereferenced on the device (tile_build.cu, pme.cu): a malformed CSR must not get that far
    MC_REQUIRE(c, start[0] == 0, "mc_set_exclusions: start[0] must be 0");
    for (int64_t i = 0; i < n; ++i) MC_REQUIRE(c, start[i + 1] >= start[i], "mc_set_exclusions: start[] must be non-decreasing");
    for (int64_t k = 0; k < m; ++k) MC_REQUIRE(c, idx[k] >= 0 && idx[k] < n, "mc_set_exclusions: partner id out of range");
    MC_CUDA(c, c->excl_start.ensure((size_t)n + 1));
    MC_CUDA(c, c->excl_idx.ensure((size_t)m));
    MC_CUDA(c, cudaMemcpy(c->excl_start.p, start, (n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
    MC_CUDA(c, cudaMemcpy(c->excl_idx.p, idx, m * sizeof(int32_t), cudaMemcpyHostToDevice));
    c->have_excl = true;
    return MC_OK;
}

extern "C" int mc_set_pairs14(mc_ctx *c, int64_t m, const int32_t *pairs, float scale_lj, float scale_q) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    cudaSetDevice(c->device);
    c->forces_valid = false;
    c->scale14_lj = scale_lj;
    c->scale14_q = scale_q;
    if (m <= 0 || !pairs) { c->have_p14 = false; return MC_OK; }
    const int64_t n = c->n_global;
    // symmetric per-atom rows (original ids) so that each atom sums its own partners
    std::vector<int32_t> start((size_t)n + 1, 0), idx((size_t)2 * m);
    for (int64_t k = 0; k < m; ++k) {
        const int32_t i = pairs[2 * k], j = pairs[2 * k + 1];
        MC_REQUIRE(c, i >= 0 && j >= 0 && i < n && j < n && i != j, "mc_set_pairs14: atom id out of range");
        start[(size_t)i + 1]++; start[(size_t)j + 1]++;
    }
    for (int64_t i = 0; i < n; ++i) start[(size_t)i + 1] += start[(size_t)i];
    std::vector<int32_t> cur(start.begin(), start.end() - 1);
    for (int64_t k = 0; k < m; ++k) {
        const int32_t i = pairs[2 * k], j = pairs[2 * k + 1];
        idx[(size_t)cur[(size_t)i]++] = j;
        idx[(size_t)cur[(size_t)j]++] = i;
    }
    MC_CUDA(c, c->p14_start.ensure((size_t)n + 1));
    MC_CUDA(c, c->p14_idx.ensure(idx.size()));
    MC_CUDA(c, cudaMemcpy(c->p14_start.p, start.data(), start.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    MC_CUDA(c, cudaMemcpy(c->p14_idx.p, idx.data(), idx.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    c->have_p14 = true;
    return MC_OK;
}

// ---- bonded terms ------------------------------------------------------------------
// Host layout in, device layout (int2 / int4 + parameter vectors) out; atom ids are the caller's.

template <int W, typename IdxT>
static int upload_terms(mc_ctx *c, const char *who, int64_t m, const int32_t *ids, DevBuf<IdxT> &d_ids, int *count) {
    const int64_t n = c->n_global;
    std::vector<IdxT> h((size_t)std::max<int64_t>(m, 1));
    for (int64_t k = 0; k < m; ++k) {
        int v[4] = {0, 0, 0, 0};
        for (int a = 0; a < W; ++a) {
            v[a] = ids[W * k + a];
            MC_REQUIRE(c, v[a] >= 0 && v[a] < n, std::string(who) + ": atom id out of range");
        }
        memcpy(&h[(size_t)k], v, sizeof(IdxT));
    }
    MC_CUDA(c, d_ids.ensure((size_t)std::max<int64_t>(m, 1)));
    if (m) MC_CUDA(c, cudaMemcpy(d_ids.p, h.data(), sizeof(IdxT) * (size_t)m, cudaMemcpyHostToDevice));
    *count = (int)m;
    c->forces_valid = false;
    return MC_OK;
}

extern "C" int mc_set_bonds(mc_ctx *c, int64_t m, const int32_t *pairs, const float *k_r0) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, m >= 0 && m < ((int64_t)1 << 30) && (m == 0 || (pairs && k_r0)), "mc_set_bonds: bad arguments");
    int rc = upload_terms<2, int2>(c, "mc_set_bonds", m, pairs, c->bonds, &c->n_bonds);
    if (rc != MC_OK) { c->n_bonds = 0; return rc; }
    MC_CUDA(c, c->bond_kr0.ensure((size_t)std::max<int64_t>(m, 1)));
    if (m) MC_CUDA(c, cudaMemcpy(c->bond_kr0.p, k_r0, sizeof(float2) * (size_t)m, cudaMemcpyHostToDevice));
    return MC_OK;
}

extern "C" int mc_set_angles(mc_ctx *c, int64_t m, const int32_t *triples, const float *k_theta0) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, m >= 0 && m < ((int64_t)1 << 30) && (m == 0 || (triples && k_theta0)), "mc_set_angles: bad arguments");
    int rc = upload_terms<3, int4>(c, "mc_set_angles", m, triples, c->angles, &c->n_angles);
    if (rc != MC_OK) { c->n_angles = 0; return rc; }
    MC_CUDA(c, c->angle_kt0.ensure((size_t)std::max<int64_t>(m, 1)));
    if (m) MC_CUDA(c, cudaMemcpy(c->angle_kt0.p, k_theta0, sizeof(float2) * (size_t)m, cudaMemcpyHostToDevice));
    return MC_OK;
}

extern "C" int mc_set_dihedrals(mc_ctx *c, int64_t m, const int32_t *quads, const float *pk_n_phase) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, m >= 0 && m < ((int64_t)1 << 30) && (m == 0 || (quads && pk_n_phase)), "mc_set_dihedrals: bad arguments");
    int rc = upload_terms<4, int4>(c, "mc_set_dihedrals", m, quads, c->dihedrals, &c->n_dihedrals);
    if (rc != MC_OK) { c->n_dihedrals = 0; return rc; }
    std::vector<float4> prm((size_t)std::max<int64_t>(m, 1));
    for (int64_t k = 0; k < m; ++k) prm[(size_t)k] = make_float4(pk_n_phase[3 * k], pk_n_phase[3 * k + 1], pk_n_phase[3 * k + 2], 0.f);
    MC_CUDA(c, c->dihedral_prm.ensure(prm.size()));
    if (m) MC_CUDA(c, cudaMemcpy(c->dihedral_prm.p, prm.data(), sizeof(float4) * (size_t)m, cudaMemcpyHostToDevice));
    return MC_OK;
}

extern "C" int mc_set_hbond_constraints(mc_ctx *c, int64_t m, const int32_t *clusters, const float *lengths) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, !c->comm_active, "mc_set_hbond_constraints: constraints on a decomposed handle are not supported yet");
    MC_REQUIRE(c, m >= 0 && m < ((int64_t)1 << 30) && (m == 0 || (clusters && lengths)), "mc_set_hbond_constraints: bad arguments");
    const int64_t n = c->n_global;
    std::vector<int4> h((size_t)std::max<int64_t>(m, 1));
    std::vector<uint8_t> seen((size_t)std::max<int64_t>(n, 1), 0);  // one thread owns a cluster: no atom may be in two
    int n_con = 0;
    for (int64_t k = 0; k < m; ++k) {
        const int32_t *q = clusters + 4 * k;
        for (int a = 0; a < 4; ++a) {
            if (q[a] < 0 || q[a] >= n) continue;  // range errors are reported below
            MC_REQUIRE(c, !seen[(size_t)q[a]], "mc_set_hbond_constraints: an atom appears in two clusters");
            MC_REQUIRE(c, c->h_in_water.empty() || !c->h_in_water[(size_t)q[a]],
                       "mc_set_hbond_constraints: an atom of a rigid water (mc_set_rigid_waters) cannot also sit in a hydrogen cluster");
            seen[(size_t)q[a]] = 1;
        }
        MC_REQUIRE(c, q[0] >= 0 && q[0] < n, "mc_set_hbond_constraints: heavy atom id out of range");
        for (int a = 1; a < 4; ++a) {
            MC_REQUIRE(c, q[a] >= -1 && q[a] < n && q[a] != q[0], "mc_set_hbond_constraints: hydrogen id out of range");
            if (q[a] >= 0) {
                MC_REQUIRE(c, lengths[3 * k + a - 1] > 0.f, "mc_set_hbond_constraints: constrained length must be positive");
                ++n_con;
            }
        }
        h[(size_t)k] = make_int4(q[0], q[1], q[2], q[3]);
    }
    MC_CUDA(c, c->hclusters.ensure(h.size()));
    MC_CUDA(c, c->hdist.ensure((size_t)std::max<int64_t>(3 * m, 1)));
    MC_CUDA(c, c->shake_fail.ensure(1));
    MC_CUDA(c, cudaMemset(c->shake_fail.p, 0, sizeof(int)));
    if (m) {
        MC_CUDA(c, cudaMemcpy(c->hclusters.p, h.data(), sizeof(int4) * (size_t)m, cudaMemcpyHostToDevice));
        MC_CUDA(c, cudaMemcpy(c->hdist.p, lengths, sizeof(float) * 3 * (size_t)m, cudaMemcpyHostToDevice));
    }
    c->n_hclusters = (int)m;
    c->n_hconstraints = n_con;
    if (m) c->h_in_hcluster.swap(seen); else c->h_in_hcluster.clear();
    return MC_OK;
}

extern "C" int mc_set_virtual_sites(mc_ctx *c, int64_t m, const int32_t *quads, float a, float b) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, !c->comm_active, "mc_set_virtual_sites: virtual sites on a decomposed handle are not supported yet");
    MC_REQUIRE(c, m >= 0 && m < ((int64_t)1 << 30) && (m == 0 || quads), "mc_set_virtual_sites: bad arguments");
    {
        // vsite_spread hands a site's force to its parents with plain read-modify-writes, vsite_construct writes the site: a site
        // id, or a parent, that appears in two sites would race (1 = used as a site, 2 = used as a parent)
        std::vector<uint8_t> used((size_t)std::max<int64_t>(c->n_global, 1), 0);
        for (int64_t k = 0; k < m; ++k) {
            for (int a = 0; a < 4; ++a) {
                const int32_t id = quads[4 * k + a];
                if (id < 0 || id >= c->n_global) continue;  // range errors are reported by upload_terms below
                MC_REQUIRE(c, !used[(size_t)id], a == 0 ? "mc_set_virtual_sites: a site id appears twice (or is a parent of another site)"
                                                         : "mc_set_virtual_sites: a parent atom appears in two sites (or is itself a site)");
                used[(size_t)id] = a == 0 ? 1 : 2;
            }
        }
    }
    int rc = upload_terms<4, int4>(c, "mc_set_virtual_sites", m, quads, c->vsites, &c->n_vsites);
    if (rc != MC_OK) { c->n_vsites = 0; return rc; }
    c->vsite_a = a; c->vsite_b = b;
    return MC_OK;
}

extern "C" int mc_set_thermostat(mc_ctx *c, int kind, float temperature_k, float gamma_per_ps, uint64_t seed) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    MC_REQUIRE(c, kind == MC_THERMOSTAT_NONE || kind == MC_THERMOSTAT_LANGEVIN || kind == MC_THERMOSTAT_CSVR, "mc_set_thermostat: unknown kind");
    // Langevin acts on owned rows with noise keyed by (seed, step, original atom id): the same numbers on any decomposition.
    // CSVR needs the global kinetic energy inside the step: one 24-byte all-reduce per step on the engine stream.
    MC_REQUIRE(c, kind == MC_THERMOSTAT_NONE || (temperature_k >= 0.f && gamma_per_ps >= 0.f), "mc_set_thermostat: negative temperature or friction");
    c->langevin = kind == MC_THERMOSTAT_LANGEVIN;
    c->csvr = kind == MC_THERMOSTAT_CSVR;
    c->lgv_temperature = temperature_k;
    c->lgv_gamma = gamma_per_ps;
    c->lgv_seed = seed;
    c->lgv_step = 0;
    return MC_OK;
}

extern "C" int mc_set_pme(mc_ctx *c, int k1, int k2, int k3) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, !c->comm_active, "mc_set_pme: reciprocal space on a decomposed handle is not supported yet");
    MC_REQUIRE(c, (k1 == 0 && k2 == 0 && k3 == 0) || c->periodic, "mc_set_pme: needs a periodic box");
    const char *msg = "";
    int rc = pme_configure(&c->pme, k1, k2, k3, c->st, &msg);
    if (rc != MC_OK) return fail(c, rc, std::string("mc_set_pme: ") + msg);
    c->forces_valid = false;
    return MC_OK;
}

extern "C" int mc_set_rigid_waters(mc_ctx *c, int64_t m, const int32_t *triples, float d_oh, float d_hh, float m_o, float m_h) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, !c->comm_active, "mc_set_rigid_waters: constraints on a decomposed handle are not supported yet");
    MC_REQUIRE(c, m >= 0 && m < ((int64_t)1 << 30) && (m == 0 || triples), "mc_set_rigid_waters: bad arguments");
    MC_REQUIRE(c, m == 0 || (d_oh > 0.f && d_hh > 0.f && d_hh < 2.f * d_oh && m_o > 0.f && m_h > 0.f),
               "mc_set_rigid_waters: need 0 < d_hh < 2 d_oh and positive masses");
    std::vector<uint8_t> seen((size_t)std::max<int64_t>(c->n_global, 1), 0);  // one thread owns a molecule: no atom may be in two
    for (int64_t k = 0; k < 3 * m; ++k) {
        const int32_t id = triples[k];
        if (id < 0 || id >= c->n_global) continue;  // range errors are reported by upload_terms below
        MC_REQUIRE(c, !seen[(size_t)id], "mc_set_rigid_waters: an atom appears in two waters (or twice in one)");
        MC_REQUIRE(c, c->h_in_hcluster.empty() || !c->h_in_hcluster[(size_t)id],
                   "mc_set_rigid_waters: an atom of a hydrogen cluster (mc_set_hbond_constraints) cannot also sit in a rigid water");
        seen[(size_t)id] = 1;
    }
    int rc = upload_terms<3, int4>(c, "mc_set_rigid_waters", m, triples, c->waters, &c->n_waters);
    if (rc != MC_OK) { c->n_waters = 0; c->h_in_water.clear(); return rc; }
    if (m) c->h_in_water.swap(seen); else c->h_in_water.clear();
    c->water_d_oh = d_oh; c->water_d_hh = d_hh; c->water_m_o = m_o; c->water_m_h = m_h;
    return MC_OK;
}

extern "C" int mc_set_positions(mc_ctx *c, const mc_float4 *xyzq) {
    if (!c || !xyzq) return MC_E_INVALID;
    MC_FLUSH(c);
    MC_REQUIRE(c, !c->comm_active, "mc_set_positions: not available on a decomposed handle; use mc_set_atoms");
    cudaSetDevice(c->device);
    const int64_t n = c->n;
    if (n == 0) return MC_OK;
    MC_CUDA(c, c->stage.ensure((size_t)n));
    MC_CUDA(c, cudaMemcpyAsync(c->stage.p, xyzq, n * sizeof(float4), cudaMemcpyHostToDevice, c->st));
    launch_scatter_from_orig((int)n, c->stage.p, c->orig[c->cur].p, c->xyzq[c->cur].p, 0, c->st, &c->launches);
    MC_CUDA(c, cudaStreamSynchronize(c->st));
    c->list_valid = false;
    c->forces_valid = false;
    return MC_OK;
}

extern "C" int mc_set_velocities(mc_ctx *c, const mc_float4 *vel) {
    if (!c || !vel) return MC_E_INVALID;
    MC_FLUSH(c);
    MC_REQUIRE(c, !c->comm_active, "mc_set_velocities: not available on a decomposed handle; use mc_set_atoms");
    cudaSetDevice(c->device);
    const int64_t n = c->n;
    if (n == 0) return MC_OK;
    MC_CUDA(c, c->stage.ensure((size_t)n));
    MC_CUDA(c, cudaMemcpyAsync(c->stage.p, vel, n * sizeof(float4), cudaMemcpyHostToDevice, c->st));
    launch_scatter_from_orig((int)n, c->stage.p, c->orig[c->cur].p, c->vel[c->cur].p, 0, c->st, &c->launches);
    MC_CUDA(c, cudaStreamSynchronize(c->st));
    return MC_OK;
}

extern "C" int mc_set_option(mc_ctx *c, const char *name, double value) {
    if (!c || !name) return MC_E_INVALID;
    MC_FLUSH(c);
    const std::string k(name);
    if (k == "pair_lanes") {
        const int v = (int)value;
        MC_REQUIRE(c, v == 4 || v == 8 || v == 16 || v == 32, "mc_set_option: pair_lanes must be 4, 8, 16 or 32");
        c->pair_lanes = v;
    } else if (k == "pair_uniform") {
        c->pair_uniform = value != 0.0;
        c->list_valid = false;  // the warp-uniform loop wants rows partitioned inner / skin shell: built on request only
    } else if (k == "sync_rebuild") {
        c->sync_rebuild = value != 0.0;
    } else if (k == "tile_sweep") {
        c->use_tile = value != 0.0;
        c->list_valid = false;
    } else if (k == "pair_tile_stages") {
        MC_REQUIRE(c, value == 0.0 || (value >= 2.0 && value <= 4.0), "mc_set_option: pair_tile_stages is 0 (automatic), 2, 3 or 4");
        c->pair_tile_stages = (int)value;
    } else if (k == "pair_tile") {
        c->use_pair_tile = value == 0.0 ? 0 : (value == 2.0 ? 2 : 1);
        c->pair_tile_fits = true;
        c->list_valid = false;
    } else if (k == "profiling") {
        c->profiling = value != 0.0;
    } else if (k == "profile_every") {
        c->prof_every = std::max(1, (int)value);
    } else if (k == "subcell_sort") {
        c->subcell_sort = value != 0.0;
        c->grid_dirty = true;
        c->list_valid = false;
    } else if (k == "dd_migrate") {
        MC_REQUIRE(c, c->comm != nullptr, "mc_set_option: dd_migrate needs a communicator");
        comm_set_migrate(c, value != 0.0);
    } else if (k == "halo_fused") {
        c->halo_fused = value != 0.0;
    } else if (k == "rebuild_every") {
        c->rebuild_every = (int)value;
    } else if (k == "zero_com_drift") {
        MC_REQUIRE(c, value >= 0.0 && !c->comm_active, "mc_set_option: zero_com_drift = k >= 0 steps, single-GPU handles");
        c->com_every = (int)value;
    } else if (k == "defer_tail") {
        c->defer_tail = value != 0.0;
    } else if (k == "build_variant") {
        MC_REQUIRE(c, value == 1.0 || value == 2.0, "mc_set_option: build_variant is 1 (tile_build_kernel) or 2 (rows_build_kernel)");
        c->build_variant = (int)value;
        c->list_valid = false;
    } else if (k == "rows_interleave") {
        c->rows_interleave = value != 0.0 ? 1 : 0;
        c->list_valid = false;
    } else if (k == "build_split") {
        MC_REQUIRE(c, value >= 0.0 && value <= 64.0, "mc_set_option: build_split is 0 (automatic) .. 64");
        c->build_split = (int)value;
    } else if (k == "rows_dense") {  // A/B knob: 0 = dense systems build as in round 2 (8 consumer warps, count + second sweep)
        c->rows_dense = value != 0.0;
    } else if (k == "rows_min_blocks") {  // register budget of rows_build_kernel: 3 CTAs per SM (72 registers) or 2 (112)
        MC_REQUIRE(c, value == 2.0 || value == 3.0, "mc_set_option: rows_min_blocks is 2 or 3");
        c->rows_min_blocks = (int)value;
    } else if (k == "row_stage_limit") {  // testing knob: rows longer than this take the two-sweep path of rows_build_kernel
        MC_REQUIRE(c, value >= 0.0, "mc_set_option: row_stage_limit >= 0 (0 = no limit)");
        c->row_stage_limit = (uint32_t)value;
        c->list_valid = false;
    } else if (k == "early_tail") {
        c->early_tail = value != 0.0;
    } else if (k == "lazy_sync") {
        c->lazy_sync = value != 0.0;
    } else if (k == "fused_lanes") {
        MC_REQUIRE(c, value == 0.0 || value == 8.0 || value == 16.0 || value == 32.0, "mc_set_option: fused_lanes is 0 (automatic), 8, 16 or 32");
        c->fused_lanes = (int)value;
    } else if (k == "fused_brute") {
        c->fused_brute = value != 0.0;
    } else if (k == "fused_steps") {
        c->fused_steps = value != 0.0;
    } else {
        return fail(c, MC_E_INVALID, "mc_set_option: unknown option '" + k + "'");
    }
    return MC_OK;
}

// ---- neighbour build -------------------------------------------------------------------------------

static float list_radius(const mc_ctx *c) { return std::max(c->rc_lj, c->rc_q) + c->skin; }

static int setup_grid(mc_ctx *c) {
    const float r_list = list_radius(c);
    MC_CUDA(c, c->grid.ensure(1));
    MC_CUDA(c, c->bbox.ensure(8));
    const double cw_min = (double)r_list * 1.001 + 1e-3;  // same margin as oracle/md_oracle.c
    if (c->periodic) {
        GridParams g;
        long long ncell = 1;
        for (int a = 0; a < 3; ++a) {
            if (2.0f * r_list > c->ext[a])
                return fail(c, MC_E_INVALID, "neighbour build: cutoff + skin exceeds half the periodic box on axis " +
                                                 std::to_string(a) + " (minimum image would be ambiguous)");
            int m = (int)std::floor((double)c->ext[a] / cw_min);
            m = std::max(1, std::min(m, 1024));
            g.nc[a] = m;
            g.lo[a] = c->lo[a];
            g.ext[a] = c->ext[a];
            g.inv_ext[a] = 1.0f / c->ext[a];
            g.inv_cw[a] = (float)((double)m / (double)c->ext[a]);
            ncell *= m;
        }
        g.ncell = (int)ncell;
        g.periodic = 1;
        g.z_ring = 1;
        g.kz_off = 0;
        g.ncz_global = g.nc[2];
        g.row_l0 = 0;
        g.row_l1 = g.nc[2];
        g.sub_bits = (c->subcell_sort && ncell <= (1ll << (32 - MC_SUB_BITS - 1))) ? MC_SUB_BITS : 0;
        c->ncell_cap = (size_t)ncell;
        c->h_grid = g;
        MC_CUDA(c, cudaMemcpyAsync(c->grid.p, &c->h_grid, sizeof(GridParams), cudaMemcpyHostToDevice, c->st));
    } else {
        // vacuum: the grid follows the bounding box at every build; cap the cell count
        c->ncell_cap = std::max<size_t>(4096, std::min<size_t>((size_t)c->n, (size_t)1 << 22));
        c->cw_min = (float)cw_min;
    }
    MC_CUDA(c, c->cell_start.ensure(c->ncell_cap + 2));
    int bits = 1;
    while (((size_t)1 << bits) < c->ncell_cap) ++bits;
    c->key_bits = bits + (c->periodic ? c->h_grid.sub_bits : MC_SUB_BITS);
    c->grid_dirty = false;
    return MC_OK;
}

int engine_build_list(mc_ctx *c) {
    const int n = (int)c->n;
    c->bl_valid = false;  // atoms change slots
    cudaStream_t st = c->st;
    if (c->grid_dirty) { int rc = setup_grid(c); if (rc != MC_OK) return rc; }
    if (n == 0) { c->list_valid = true; return MC_OK; }
    TimedRegion tr(c, c->build_acc);
    if (!c->periodic)
        launch_bbox(c->xyzq[c->cur].p, n, c->bbox.p, c->cw_min, (int)c->ncell_cap, c->grid.p, st, &c->launches);
    MC_CUDA(c, c->keys[0].ensure((size_t)n)); MC_CUDA(c, c->keys[1].ensure((size_t)n));
    MC_CUDA(c, c->vals[0].ensure((size_t)n)); MC_CUDA(c, c->vals[1].ensure((size_t)n));
    MC_CUDA(c, c->scratch.ensure(std::max(radix_scratch_elems((size_t)n), scan_scratch_elems((size_t)n + 1)) + 64));
    launch_wrap_key(c->xyzq[c->cur].p, n, c->grid.p, c->keys[0].p, c->vals[0].p, st, &c->launches);
    uint32_t *kk[2] = {c->keys[0].p, c->keys[1].p}, *vv[2] = {c->vals[0].p, c->vals[1].p};
    const int which = radix_sort_pairs(kk, vv, (size_t)n, c->key_bits, c->scratch.p, st, &c->launches);
    const int nx = c->cur ^ 1;
    ReorderArrays ra;
    ra.xyzq_in = c->xyzq[c->cur].p; ra.xyzq_out = c->xyzq[nx].p; ra.xref = c->xref.p;
    ra.vel_in = c->vel[c->cur].p; ra.vel_out = c->vel[nx].p;
    ra.type_in = c->type[c->cur].p; ra.type_out = c->type[nx].p;
    ra.flags_in = c->flags[c->cur].p; ra.flags_out = c->flags[nx].p;
    ra.orig_in = c->orig[c->cur].p; ra.orig_out = c->orig[nx].p; ra.slot_of_orig = c->slot_of_orig.p;
    ra.cell_start = c->cell_start.p;
    ra.mark_interior = (c->skin < 0.5f * list_radius(c)) ? 1 : 0;
    launch_reorder(n, kk[which], vv[which], c->grid.p, ra, st, &c->launches);
    c->cell_of_slot = kk[which];  // sorted keys = cell of every slot, valid until the next sort
    c->cur = nx;
    c->identity_order = false;
    tr.stop();
    return engine_build_rows(c);
}

// Verlet rows for the atoms of the row layers, from cell-ordered arrays + cell_start (shared by the
// single-GPU path above and the decomposed path in comm.cu).
int engine_build_rows(mc_ctx *c) {
    const int n = (int)c->n;
    cudaStream_t st = c->st;
    TimedRegion tr(c, c->build_acc);
    const float r_list = list_radius(c);
    const float rl2 = r_list * r_list;
    const int n_rows = n;  // every local slot may carry a row; ghost layers are skipped by the grid's row_l0/row_l1
    MC_CUDA(c, cudaMemsetAsync(c->nbr_count.p, 0, sizeof(uint32_t) * (size_t)n, st));
    const int32_t *es = c->have_excl ? c->excl_start.p : nullptr, *ei = c->have_excl ? c->excl_idx.p : nullptr;
    uint32_t *h_ctl = reinterpret_cast<uint32_t *>(c->h_pinned);
    size_t total = 0;
    bool tiled = c->use_tile;
    bool ilv = false, no_row_hint = false;
    c->ilv_valid = false;
    const float rc_in = std::max(c->rc_lj, c->rc_q);
    const float rc2_inner = rc_in * rc_in;
    const int grid_cells = c->periodic ? c->h_grid.ncell : (int)c->ncell_cap;
    const int est_cells = c->periodic ? c->h_grid.ncell : std::max(1, n / 256);
    int split = std::max(1, std::min(8, (4 * c->n_sms + est_cells - 1) / est_cells));
    if (c->build_split > 0) split = c->build_split;  // option "build_split" (A/B): slices per cell of the list build
    // compact rows (16-bit tile-local indices) for the TMA-staged force kernel whenever its tile + LJ table fit shared memory
    bool compact = tiled && c->periodic && c->pair_tile_fits && (c->use_pair_tile == 1 || (c->use_pair_tile == 2 && c->n_rows_sorted() >= 16384));
    while (tiled) {
        // single-pass TMA-staged build (tile_build.cu); tile and list capacities adapt on demand
        MC_CUDA(c, c->tile_need.ensure(8));
        if (compact) { if (!c->nbr_list16.p) MC_CUDA(c, c->nbr_list16.ensure(1024)); }
        else if (!c->nbr_list.p) MC_CUDA(c, c->nbr_list.ensure(1024));
        const size_t cap_now = compact ? c->nbr_list16.n : c->nbr_list.n;
        const bool rows_v2 = c->build_variant == 2 && !compact && !c->pair_uniform;
        if (rows_v2) MC_CUDA(c, c->rows_plan.ensure(rows_plan_words(grid_cells)));
        // rows_build_kernel: a tile capacity close to the largest tile of the previous build leaves room for more tiles in flight
        uint32_t tile_cap_now = c->tile_cap;
        if (rows_v2 && c->tile_max_m) tile_cap_now = std::min(c->tile_cap, (c->tile_max_m + c->tile_max_m / 8u + 63u) & ~31u);
        launch_tile_build(n_rows, grid_cells, split, c->n_sms, c->xyzq[c->cur].p, c->cell_start.p, c->grid.p, rl2, rc2_inner,
                          c->orig[c->cur].p, es, ei, c->nbr_count.p, c->nbr_start.p,
                          compact ? static_cast<void *>(c->nbr_list16.p) : static_cast<void *>(c->nbr_list.p), compact, c->pair_uniform,
                          (uint32_t)std::min<size_t>(cap_now, 0xffffffffu), tile_cap_now, c->tile_need.p, st, &c->launches,
                          c->build_variant, no_row_hint ? 0u : (c->row_stage_limit ? std::min(c->row_len_hint, c->row_stage_limit) : c->row_len_hint),
                          rows_v2 ? c->rows_plan.p : nullptr, c->rows_dense ? c->rows_min_blocks : -c->rows_min_blocks,
                          c->rows_dense ? c->slot_of_orig.p : nullptr);
        // quad-interleaved copy for the 8-lane force kernel: its sizes pass rides in front of the build's own host sync
        // (ctl[7] = entries of the copy), the copy itself follows once the list is known to be complete
        ilv = !compact && !c->pair_uniform && c->pair_lanes == 8 && c->rows_interleave;
        if (ilv) {
            MC_CUDA(c, c->ilv_qbase.ensure((size_t)(n_rows + 3) / 4 + 1));
            launch_rows_interleave_sizes(n_rows, c->nbr_count.p, c->ilv_qbase.p, c->tile_need.p + 7, st, &c->launches);
        }
        MC_CUDA(c, cudaMemcpyAsync(h_ctl, c->tile_need.p, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        MC_CUDA(c, cudaStreamSynchronize(st));
        if ((h_ctl[3] & 2u) && !(h_ctl[3] & 1u)) {
            // single-sweep build of a dense system: a row outgrew the space claimed from the previous build's longest row
            // (+ 25 %); the sweep counted on, so the new longest row is known -- build again with it
            // -- this time counting first (always fits), the next build of the system uses the new length
            c->row_len_hint = std::max(c->row_len_hint, h_ctl[6]);
            no_row_hint = true;
            continue;
        }
        if (h_ctl[3] != 0) {  // a neighbourhood did not fit the tile
            uint32_t need = (h_ctl[2] + h_ctl[2] / 4 + 127u) & ~31u;  // 25 % head-room + NaN padding to whole chunks
            const uint32_t max_atoms = (rows_v2 && c->rows_dense) ? tile_sweep_max_atoms_dense() : tile_sweep_max_atoms();
            if (need > max_atoms && ((h_ctl[2] + 127u) & ~31u) <= max_atoms) need = max_atoms;
            if (need <= max_atoms) { c->tile_cap = std::max(c->tile_cap, need); c->tile_max_m = 0; continue; }
            tiled = c->use_tile = false;  // too dense for shared memory: two-pass global sweep from now on
            compact = false;
            break;
        }
        total = h_ctl[1];
        if (total > cap_now) {  // the cursor ran past the list: grow it and build again
            if (compact) MC_CUDA(c, c->nbr_list16.ensure(total + total / 8 + 1024));
            else MC_CUDA(c, c->nbr_list.ensure(total + total / 8 + 1024));
            continue;
        }
        c->tile_max_m = (h_ctl[2] + 31u) & ~31u;
        c->rows_max_entries = h_ctl[4];
        if (h_ctl[6]) c->row_len_hint = h_ctl[6];  // longest row: sizes the row staging of the next build (rows_build_kernel)
        if (ilv && (uint64_t)h_ctl[7] + 4096u < 0xffffffffull) {
            MC_CUDA(c, c->nbr_ilv.ensure((size_t)h_ctl[7] + 64));
            launch_rows_interleave_copy(n_rows, c->nbr_start.p, c->nbr_count.p, c->nbr_list.p, c->ilv_qbase.p, c->nbr_ilv.p, st, &c->launches);
            c->ilv_valid = true;
        }
        // the TMA-staged force kernel wants every cell's rows as one block of <= 32 rows next to the tile in shared memory:
        // a system that is too dense for that (seen only now) is built again with global-slot rows, and stays that way
        if (compact && (h_ctl[5] > 32u || pair_tile_smem(c->tile_max_m, c->rows_max_entries, c->n_types, c->n_types > 1, nullptr, nullptr) == 0)) {
            compact = false;
            c->pair_tile_fits = false;
            continue;
        }
        if (compact) {
            // per-cell staging records of the force kernel's producer (fixed until the next build)
            uint32_t rows_cap = 0;
            pair_tile_smem(c->tile_max_m, c->rows_max_entries, c->n_types, c->n_types > 1, nullptr, &rows_cap);
            MC_CUDA(c, c->cell_plan.ensure((size_t)c->h_grid.ncell * pair_tile_plan_words()));
            MC_CUDA(c, c->cell_rowtab.ensure((size_t)c->h_grid.ncell * 32));
            launch_cell_plan(c->h_grid.ncell, c->cell_start.p, c->grid.p, c->nbr_start.p, c->nbr_count.p, c->tile_max_m, rows_cap,
                             c->cell_plan.p, c->cell_rowtab.p, c->tile_need.p, st, &c->launches);
        }
        break;
    }
    if (!tiled) {
        launch_sweep(false, n_rows, c->xyzq[c->cur].p, c->cell_start.p, c->grid.p, rl2, c->cell_of_slot, c->orig[c->cur].p, es, ei,
                     c->nbr_count.p, nullptr, nullptr, st, &c->launches);
        exclusive_scan_u32(c->nbr_count.p, c->nbr_start.p, (size_t)n_rows, 1, c->scratch.p, st, &c->launches);
        MC_CUDA(c, cudaMemcpyAsync(h_ctl, c->nbr_start.p + n_rows, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        MC_CUDA(c, cudaStreamSynchronize(st));
        total = h_ctl[0];
        if (total > c->nbr_list.n) MC_CUDA(c, c->nbr_list.ensure(total + total / 8 + 1024));
        launch_sweep(true, n_rows, c->xyzq[c->cur].p, c->cell_start.p, c->grid.p, rl2, c->cell_of_slot, c->orig[c->cur].p, es, ei,
                     c->nbr_count.p, c->nbr_start.p, c->nbr_list.p, st, &c->launches);
    }
    MC_CUDA(c, cudaMemsetAsync(c->rebuild_flag.p, 0, sizeof(int), st));
    tr.stop();
    MC_CUDA(c, cudaGetLastError());
    c->n_padded_entries = (int64_t)total;
    c->list_compact = tiled && compact;
    c->list32_valid = !c->list_compact;
    c->list_valid = true;
    c->forces_valid = false;
    c->n_rebuilds++;
    c->tail_use_split = false;  // a halo descriptor kept for a deferred force evaluation describes the old layout
    c->drift_event_valid = false;  // the atoms sit in other arrays now, written by kernels behind that event
    c->steps_since_build = 0;
    c->pairs_dirty = true;
    return MC_OK;
}

int engine_ensure_list32(mc_ctx *c) {
    if (c->list32_valid || !c->list_valid) return MC_OK;
    MC_CUDA(c, c->nbr_list.ensure((size_t)std::max<int64_t>(c->n_padded_entries, 1)));
    launch_expand_rows(c->periodic ? c->h_grid.ncell : (int)c->ncell_cap, c->cell_start.p, c->grid.p, c->nbr_start.p, c->nbr_count.p,
                       c->nbr_list16.p, c->nbr_list.p, c->st, &c->launches);
    MC_CUDA(c, cudaGetLastError());
    c->list32_valid = true;
    return MC_OK;
}

extern "C" int mc_build_neighbors(mc_ctx *c) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, c->rc_lj > 0.f, "mc_build_neighbors: call mc_set_cutoffs first");
    MC_REQUIRE(c, c->n_types > 0, "mc_build_neighbors: call mc_set_lj_table first");
    if (c->comm_active) {
        int rc = comm_rebuild(c);
        if (rc != MC_OK) return rc;
        MC_CUDA(c, cudaStreamSynchronize(c->st));
        c->collect_timings();
        return MC_OK;
    }
    int rc = engine_build_list(c);
    if (rc != MC_OK) return rc;
    MC_CUDA(c, cudaStreamSynchronize(c->st));
    c->collect_timings();
    return MC_OK;
}

// ---- forces ------------------------------------------------------------------------------------------

static NbParams make_params(const mc_ctx *c) {
    NbParams p;
    for (int a = 0; a < 3; ++a) {
        p.ext[a] = c->periodic ? c->ext[a] : 1.f;
        p.inv_ext[a] = c->periodic ? 1.0f / c->ext[a] : 1.f;
    }
    p.rc2_lj = c->rc_lj * c->rc_lj;
    p.rc2_q = c->rc_q * c->rc_q;
    p.alpha = c->alpha;
    p.sig2 = c->sig2_0;
    p.eps24 = c->eps24_0;
    p.periodic = c->periodic ? 1 : 0;
    p.n_types = c->n_types;
    return p;
}

int engine_launch_forces(mc_ctx *c, bool want_energy, const HaloSplit *hs) {
    PairLaunch L;
    L.n_rows = (int)c->n_rows_sorted();
    L.row0 = (int)c->row0;
    L.xyzq = c->xyzq[c->cur].p;
    L.type = c->type[c->cur].p;
    L.flags = c->flags[c->cur].p;
    L.energy = want_energy;
    L.uniform = c->pair_uniform;
    L.nbr_start = c->nbr_start.p; L.nbr_count = c->nbr_count.p; L.nbr_list = c->nbr_list.p;
    L.ljtab = c->ljtab.p;
    L.p = make_params(c);
    L.lj_on = c->lj_disabled ? 0 : 1;
    L.coul = c->coul_disabled ? MC_COULOMB_NONE : c->coul_mode;
    L.multi = c->n_types > 1;
    L.lanes = c->pair_lanes;
    L.force = c->force.p;
    if (c->ilv_valid && c->pair_lanes == 8 && !c->pair_uniform) { L.ilv_qbase = c->ilv_qbase.p; L.ilv_list = c->nbr_ilv.p; }
    if (c->list_compact) {
        // TMA-staged kernel over the compact rows (pair_tile.cu); a decomposed rank hands it the ready flags to wait on
        PairTileLaunch T;
        T.grid_cells = c->h_grid.ncell;
        T.n_sms = c->n_sms;
        T.xyzq = L.xyzq; T.type = L.type; T.grid = c->grid.p;
        T.plan = c->cell_plan.p; T.rowtab = c->cell_rowtab.p;
        T.list16 = c->nbr_list16.p; T.ljtab = L.ljtab;
        T.p = L.p; T.lj_on = L.lj_on; T.coul = L.coul; T.multi = L.multi; T.energy = L.energy; T.force = L.force;
        T.tile_cap = c->tile_max_m;
        T.rows_max_entries = c->rows_max_entries;
        T.force_stages = c->pair_tile_stages;
        if (hs) T.wait = hs->wait;
        TimedRegion tr(c, c->pair_acc, true);
        launch_pair_tile(T, c->st, &c->launches);
        tr.stop();
    } else if (hs) {
        TimedRegion tr(c, c->pair_acc, true);
        L.n_interior = hs->last_begin - hs->n_first;
        L.n_first = hs->n_first;
        L.wait = hs->wait;
        launch_pair_force(L, c->st, &c->launches);
        tr.stop();
    } else {
        TimedRegion tr(c, c->pair_acc, true);
        launch_pair_force(L, c->st, &c->launches);
        tr.stop();
    }
    if (c->have_p14)
        launch_pairs14(L.n_rows, L.row0, L.xyzq, L.type, c->orig[c->cur].p, c->slot_of_orig.p, c->p14_start.p, c->p14_idx.p,
                       c->ljtab.p, L.p, c->scale14_lj, c->scale14_q, L.lj_on, (L.coul != MC_COULOMB_NONE) ? 1 : 0,
                       c->force.p, c->st, &c->launches);
    if (c->n_bonds + c->n_angles + c->n_dihedrals > 0) {
        BondedTerms t;
        t.n_bonds = c->n_bonds; t.n_angles = c->n_angles; t.n_dihedrals = c->n_dihedrals;
        t.bonds = c->bonds.p; t.bond_kr0 = c->bond_kr0.p;
        t.angles = c->angles.p; t.angle_kt0 = c->angle_kt0.p;
        t.dihedrals = c->dihedrals.p; t.dihedral_prm = c->dihedral_prm.p;
        MC_CUDA(c, c->bonded_e.ensure(4));
        if (c->comm_active) {
            // every rank holds the whole term list and evaluates the terms that touch its owned rows (bonded.cuh)
            if (!c->bonded_missing.p) {
                MC_CUDA(c, c->bonded_missing.ensure(1));
                MC_CUDA(c, cudaMemsetAsync(c->bonded_missing.p, 0, sizeof(int), c->st));
            }
            t.own0 = (int)c->row0; t.own1 = (int)(c->row0 + c->n_rows_sorted());
            t.missing = c->bonded_missing.p;
        }
        launch_bonded(t, c->slot_of_orig.p, L.xyzq, L.p, c->force.p, c->bonded_e.p, want_energy, c->st, &c->launches);
    }
    if (c->pme.planned && c->periodic && !c->comm_active && L.coul == MC_COULOMB_ERFC) {
        const char *msg = "";
        int rc = pme_launch(&c->pme, (int)c->n, L.xyzq, c->lo, c->ext, c->alpha, c->force.p, want_energy, c->st, &c->launches, &msg);
        if (rc != MC_OK) return fail(c, rc, std::string("SPME: ") + msg);
        if (c->have_excl)
            pme_launch_exclusions(&c->pme, (int)c->n, L.xyzq, c->orig[c->cur].p, c->slot_of_orig.p, c->excl_start.p, c->excl_idx.p, L.p,
                                  c->force.p, want_energy, c->st, &c->launches);
    }
    if (c->n_vsites > 0)
        launch_vsite_spread(c->n_vsites, c->vsites.p, c->slot_of_orig.p, c->force.p, c->vsite_a, c->vsite_b, c->st, &c->launches);
    MC_CUDA(c, cudaGetLastError());
    c->forces_valid = true;
    c->forces_have_energy = want_energy;
    return MC_OK;
}

static int ensure_ready(mc_ctx *c, const char *who) {
    MC_REQUIRE(c, c->rc_lj > 0.f, std::string(who) + ": call mc_set_cutoffs first");
    MC_REQUIRE(c, c->n_types > 0, std::string(who) + ": call mc_set_lj_table first");
    if (!c->list_valid) {
        int rc = c->comm_active ? comm_rebuild(c) : engine_build_list(c);
        if (rc != MC_OK) return rc;
    }
    return MC_OK;
}

// mc_step with external forces may return with the last step's force evaluation and second half kick still open
// (see there).  Every entry point that observes or changes anything but positions closes them first.
// Closes the step a pipelined mc_step left open: (scheduled rebuild of a decomposed run,) force evaluation of the positions
// reached, second half kick with THAT call's external forces.  Nothing here synchronises.
static int collect_pending_epilogue(mc_ctx *c, int *warn);

// First half of closing the open step: (rebuild when one is due,) force evaluation of the positions reached.  A pipelined call
// launches it itself right before it returns (option early_tail), so that it runs while the caller is away -- reading the
// snapshot, preparing the next array -- instead of waiting for the next call to launch it; whoever closes the step finds the
// forces valid and goes straight to the half kick (and must not rebuild in between: the forces are in the order they were
// computed in).
static int tail_forces(mc_ctx *c) {
    if (c->forces_valid && !c->tail_rebuild) return MC_OK;
    int rc;
    bool fresh_ghosts = false;
    if (c->tail_rebuild) {
        if ((rc = comm_rebuild(c)) != MC_OK) return rc;  // every rank of the run is here: same schedule
        c->tail_rebuild = false;
        fresh_ghosts = true;
    } else if ((rc = ensure_ready(c, "mc_step (deferred half kick)")) != MC_OK) {
        return rc;
    }
    if (!c->forces_valid) {
        const bool use_split = c->tail_use_split && !fresh_ghosts && c->list_valid;
        if (c->comm_active && !use_split && !fresh_ghosts && (rc = comm_halo_positions(c)) != MC_OK) return rc;
        if ((rc = engine_launch_forces(c, false, use_split ? &c->tail_split : nullptr)) != MC_OK) return rc;
    }
    c->tail_use_split = false;
    return MC_OK;
}

static int close_tail(mc_ctx *c) {
    int rc = tail_forces(c);
    if (rc != MC_OK) return rc;
    const size_t r0 = (size_t)c->row0;
    launch_kick_drift((int)c->n_rows_sorted(), c->xyzq[c->cur].p + r0, c->vel[c->cur].p + r0, c->force.p + r0, c->tail_ext,
                      c->orig[c->cur].p + r0, c->flags[c->cur].p + r0, c->xref.p + r0, 0.5f * c->tail_dt, 0.f, 0.f, 0.f,
                      c->rebuild_flag.p, c->st, &c->launches);
    c->tail_pending = false;
    return MC_OK;
}

int engine_flush_tail(mc_ctx *c) {
    if (!c->tail_pending) return MC_OK;
    cudaSetDevice(c->device);
    int rc = collect_pending_epilogue(c, nullptr);  // (a pipelined call returns before its kernels have finished)
    if (rc != MC_OK) return rc;
    rc = close_tail(c);
    if (rc != MC_OK) return rc;
    MC_CUDA(c, cudaStreamSynchronize(c->st));
    return MC_OK;
}

extern "C" int mc_compute_forces(mc_ctx *c) {
    if (!c) return MC_E_INVALID;
    cudaSetDevice(c->device);
    MC_FLUSH_OBS(c);
    c->prof_now = true;
    int rc = ensure_ready(c, "mc_compute_forces");
    if (rc != MC_OK) return rc;
    if (c->comm_active && (rc = comm_halo_positions(c)) != MC_OK) return rc;
    if ((rc = engine_launch_forces(c, true)) != MC_OK) return rc;
    MC_CUDA(c, cudaStreamSynchronize(c->st));
    c->collect_timings();
    return MC_OK;
}

// ---- velocity Verlet -----------------------------------------------------------------------------------

// Spin on a pinned host word until kick_drift's last block has published the expected step tag.
static int wait_flag_tag(mc_ctx *c, volatile int *word, int tag) {
    for (long spins = 0;; ++spins) {
        if ((*word >> 2) == tag) return MC_OK;
        if ((spins & 0xfffff) == 0xfffff) {  // every ~1M polls: has the stream died?
            cudaError_t e = cudaStreamQuery(c->st);
            if (e != cudaSuccess && e != cudaErrorNotReady) return fail(c, MC_E_CUDA, std::string("mc_step: ") + cudaGetErrorString(e));
            if (e == cudaSuccess && (*word >> 2) != tag) return fail(c, MC_E_CUDA, "mc_step: rebuild flag never published");
        }
    }
}

static int apply_barostat(mc_ctx *c, float dt);

// What follows the one synchronisation of an mc_step call: flag words, timing, warnings.  Run at the end of the call, or -- for a
// pipelined call that returned early (StepEpilogue::pending) -- by the next call / by whoever closes the open step.
static int step_epilogue(mc_ctx *c, const StepEpilogue &E) {
    cudaStream_t st = c->st;
    int *h_flag = reinterpret_cast<int *>(c->h_pinned) + 8;
    int *h_agree = reinterpret_cast<int *>(c->h_pinned) + 12;
    const int n_steps = E.n_steps;
    MC_CUDA(c, cudaStreamSynchronize(st));
    if (E.trace_dev) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, c->trace_ev[0], c->trace_ev[1]) == cudaSuccess) c->trace_dev[0] += t;
        if (cudaEventElapsedTime(&t, c->trace_ev[2], c->trace_ev[3]) == cudaSuccess) c->trace_dev[1] += t;
        if (cudaEventElapsedTime(&t, c->trace_ev[4], c->trace_ev[5]) == cudaSuccess) c->trace_dev[2] += t;
        if (cudaEventElapsedTime(&t, c->trace_ev[5], c->ev_step_b) == cudaSuccess) c->trace_dev[3] += t;
    }
    float ms = 0.f;
    MC_CUDA(c, cudaEventElapsedTime(&ms, c->ev_step_a, c->ev_step_b));
    c->last_step_ms = ms;
    // the flag of the last drift has not been acted upon: make the next evaluation rebuild first
    if (E.pipelined && n_steps > 0 && !E.skip_prev && (h_flag[(n_steps - 1) & 1] & 3) != 0) c->list_valid = false;
    if (E.pipelined && n_steps > 0 && (h_flag[(n_steps - 1) & 1] & 2))
        return fail(c, MC_E_INVALID, "mc_step: non-finite coordinates (the simulation blew up)");
    bool stale_list = false;
    if (E.flags_arrive) {
        // the previous call's flags, the same words on every rank: maximum over the ranks, then as below
        h_agree[0] = h_agree[1] = 0;
        for (int r = 0; r < E.n_ranks_f; ++r) {
            h_agree[0] = std::max(h_agree[0], c->h_flags_all[2 * r]);
            h_agree[1] |= c->h_flags_all[2 * r + 1];
        }
        // (a rebuild between the two calls has cleared the displacement word: nothing to report then)
    }
    if (E.check_flag || E.flags_arrive) {
        if (h_agree[1] & MC_HALO_ERR_TIMEOUT)
            return fail(c, MC_E_COMM, "mc_step: a neighbour rank did not signal its halo push within 2 s (peer died or ranks out of step)");
        if (h_agree[0] & 2) return fail(c, MC_E_INVALID, "mc_step: non-finite coordinates (the simulation blew up)");
        // An atom moved more than skin/2 between two builds: the schedule was too long for this system (sudden heating, a
        // caller-chosen rebuild_every).  The list is rebuilt before the next evaluation -- on every rank of a decomposed
        // run, which all see the same reduced flag -- the adaptive interval is halved, and the caller is told
        // (MC_W_STALE_LIST: pairs inside the cutoff may have been missing from the last steps' forces).
        if (h_agree[0] != 0) {
            c->n_list_violations++;
            c->list_valid = false;
            stale_list = true;
            if (c->comm_active) comm_shrink_interval(c);
        }
    }
    if (c->comm_active && c->bonded_missing.p && n_steps > 0) {
        int bad = 0;
        MC_CUDA(c, cudaMemcpy(&bad, c->bonded_missing.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (bad) {
            MC_CUDA(c, cudaMemset(c->bonded_missing.p, 0, sizeof(int)));
            return fail(c, MC_E_INVALID, "mc_step: a bonded term reaches beyond this rank's ghost layer (bonded partners must lie within cutoff + skin)");
        }
    }
    if (c->n_hclusters > 0 && n_steps > 0) {
        int bad = 0;
        MC_CUDA(c, cudaMemcpy(&bad, c->shake_fail.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (bad) {
            MC_CUDA(c, cudaMemset(c->shake_fail.p, 0, sizeof(int)));
            return fail(c, MC_E_INVALID, "mc_step: SHAKE did not converge for " + std::to_string(bad) + " hydrogen-bond cluster steps (time step too long?)");
        }
    }
    c->collect_timings();
    if (stale_list) {
        c->err = "mc_step: an atom moved more than skin/2 between two list builds; the list is rebuilt before the next evaluation";
        return MC_W_STALE_LIST;
    }
    return MC_OK;
}

// the epilogue a pipelined call left behind (errors are returned; a warning is kept in *warn)
static int collect_pending_epilogue(mc_ctx *c, int *warn) {
    if (!c->epi.pending) return MC_OK;
    c->epi.pending = false;
    const int rc = step_epilogue(c, c->epi);
    if (rc > 0) { if (warn) *warn = rc; return MC_OK; }
    return rc;
}

extern "C" int mc_step(mc_ctx *c, float dt, int n_steps, const float *ext_forces) {
    if (!c) return MC_E_INVALID;
    cudaSetDevice(c->device);
    MC_REQUIRE(c, n_steps >= 0 && dt > 0.f, "mc_step: n_steps >= 0 and dt > 0 required");
    if (!c->comm_active && c->n_global == 0) {  // an empty system steps trivially (no block would publish the rebuild flag)
        c->n_steps += n_steps;
        return MC_OK;
    }
    c->prof_now = true;
    cudaStream_t st = c->st;
    struct Tr {  // host-side phase clock (MC_TRACE_STEP)
        mc_ctx *c; bool on; std::chrono::steady_clock::time_point t;
        void lap(int k) {
            if (!on) return;
            const auto now = std::chrono::steady_clock::now();
            const double d = std::chrono::duration<double>(now - t).count();
            c->trace_t[k] += d;
            if (d > 5e-3) fprintf(stderr, "[mc_step stall, device %d] call %lld: phase %d took %.2f ms (steps since build %d)\n", c->device,
                                  (long long)c->trace_calls, k, d * 1e3, c->steps_since_build);
            t = now;
        }
    } trc{c, c->trace_step && n_steps == 1 && ext_forces != nullptr /* the per-step calls of an end-to-end loop */, std::chrono::steady_clock::now()};
    if (trc.on) {
        c->trace_calls++;
        if (!c->trace_ev[0]) for (int k = 0; k < 8; ++k) cudaEventCreate(&c->trace_ev[k]);
    }
    const bool trace_dev = trc.on && c->trace_ev[7] != nullptr;
    // decomposed runs rebuild on a schedule every rank derives from the same numbers (no per-step agreement)
    if (c->comm_active && n_steps > 0) comm_first_interval(c, dt);
    const bool pipelined = !c->comm_active && c->rebuild_every <= 0 && !c->sync_rebuild;
    // External forces are the one per-call input of a step (MdState::step(dev, dt, Some(forces)), reference
    // src/mol_alignment.rs:346).  Their upload must not sit in front of the kernels: it goes to a stream of its own
    // (alternating device buffers), and the call returns after the LAST DRIFT -- the positions the caller reads back --
    // leaving that step's force evaluation and second half kick open (`tail_pending`).  The next call starts its
    // upload, runs the open force evaluation meanwhile (it does not depend on the new array), closes the half kick
    // with the PREVIOUS array and only then waits for the upload.  Everything that observes more than positions
    // closes the tail first (engine_flush_tail), so the deferral is invisible through the ABI.
    const bool baro = c->baro_kind != MC_BAROSTAT_NONE && c->periodic && !c->comm_active;
    // (a decomposed run rebuilds on its schedule, known to every rank: nothing to pipeline there, the deferral works as is)
    // Constrained systems (rigid waters, bonds to hydrogen) close every step on its own -- half kick, then the velocity stage
    // of RATTLE -- instead of merging the closing half kick with the next step's opening one.
    const bool constrained = c->n_waters > 0 || c->n_hclusters > 0;
    const bool defer = c->defer_tail && ext_forces != nullptr && (pipelined || (c->comm_active && !c->sync_rebuild)) && n_steps > 0 && !baro &&
                       !constrained;
    struct UploadGuard {  // whatever path leaves this function, the caller's array is no longer being read
        cudaStream_t s = nullptr;
        ~UploadGuard() { if (s) cudaStreamSynchronize(s); }
    } upload_guard;
    const float *d_ext = nullptr;
    bool wait_upload = false;
    size_t gather_chunk = 0;  // decomposed: floats per rank of the all-gather that completes the array on every rank
    if (ext_forces) {
        DevBuf<float> &buf = c->ext_k ? c->ext_force2 : c->ext_force;
        c->ext_k ^= 1;
        // A decomposed rank uploads only ITS 1/N of the caller's array (a contiguous block of original ids) over its own
        // PCIe link; the blocks are then exchanged over NVLink (one in-place ncclAllGather on the engine stream), so the
        // host-to-device traffic per rank shrinks with N instead of every rank pulling the whole array.
        int rank = 0, n_ranks = 1;
        if (c->comm_active) comm_rank_size(c, &rank, &n_ranks);
        const size_t total = (size_t)3 * c->n_global;
        const size_t chunk = c->comm_active ? (total + n_ranks - 1) / n_ranks : total;
        const size_t lo = std::min(total, (size_t)rank * chunk), cnt = std::min(chunk, total - lo);
        MC_CUDA(c, buf.ensure(chunk * (size_t)n_ranks));
        if (c->comm_active) gather_chunk = chunk;
        c->ext_upload_bytes = (int64_t)(cnt * sizeof(float));
        if (defer) {
            if (!c->st_up) {
                MC_CUDA(c, cudaStreamCreateWithFlags(&c->st_up, cudaStreamNonBlocking));
                MC_CUDA(c, cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming));
            }
            upload_guard.s = c->st_up;
            if (trace_dev) cudaEventRecord(c->trace_ev[0], c->st_up);
            if (cnt) MC_CUDA(c, cudaMemcpyAsync(buf.p + lo, ext_forces + lo, sizeof(float) * cnt, cudaMemcpyHostToDevice, c->st_up));
            if (trace_dev) cudaEventRecord(c->trace_ev[1], c->st_up);
            MC_CUDA(c, cudaEventRecord(c->ev_up, c->st_up));
            wait_upload = true;
        } else if (cnt) {
            MC_CUDA(c, cudaMemcpyAsync(buf.p + lo, ext_forces + lo, sizeof(float) * cnt, cudaMemcpyHostToDevice, st));
        }
        d_ext = buf.p;
    } else {
        c->ext_upload_bytes = 0;
    }
    int rc;
    c->drift_event_valid = false;  // (set again by a pipelined call right after its last drift)
    int lazy_rc = MC_OK;  // warning of the previous (pipelined) call's epilogue, handed back by this call
    if ((rc = collect_pending_epilogue(c, &lazy_rc)) != MC_OK) return rc;  // (the upload above is already on its way)
    trc.lap(0);
    if (trace_dev) cudaEventRecord(c->trace_ev[2], st);
    // Small plain-NVE systems: all n_steps in ONE cooperative launch (md_fused.cu) instead of 2-4 launches + a host poll per
    // step.  Up to md_fused_brute_max_atoms() atoms the kernel keeps a private all-pairs Verlet list and rebuilds it inside the
    // launch: the host's list (sort, cells, tiles) is then not needed for stepping at all.  Larger systems (up to
    // md_fused_max_atoms()) use the host's rows: the launch stops after the drift of a step that trips the displacement
    // criterion; list rebuild and force evaluation happen here, then the remaining steps go out in the next launch.
    const bool fused_ok = c->fused_steps && !c->comm_active && c->rebuild_every <= 0 && !c->sync_rebuild && !constrained && !baro && !defer &&
                          !c->langevin && !c->csvr && !c->pme.planned && c->n_vsites == 0 && c->com_every == 0 && n_steps > 0 &&
                          c->n_global <= md_fused_max_atoms() && !c->profiling;
    const bool brute = fused_ok && c->fused_brute && c->n_global <= md_fused_brute_max_atoms();
    if (!brute || c->tail_pending) c->bl_valid = false;  // any other path moves the atoms behind the private list's back
    if (c->tail_pending) {
        // the step the previous call left open: its force evaluation (it does not depend on the new array) runs under the
        // upload started above, then its second half kick with THAT call's external forces
        if ((rc = close_tail(c)) != MC_OK) return rc;
    } else if (brute) {
        MC_REQUIRE(c, c->rc_lj > 0.f, "mc_step: call mc_set_cutoffs first");
        MC_REQUIRE(c, c->n_types > 0, "mc_step: call mc_set_lj_table first");
        if (c->periodic)
            for (int a = 0; a < 3; ++a)
                MC_REQUIRE(c, 2.0f * list_radius(c) <= c->ext[a], "mc_step: cutoff + skin exceeds half the periodic box (minimum image would be ambiguous)");
    } else {
        if ((rc = ensure_ready(c, "mc_step")) != MC_OK) return rc;
        if (!c->forces_valid) {
            if (c->comm_active && (rc = comm_halo_positions(c)) != MC_OK) return rc;
            if ((rc = engine_launch_forces(c, false)) != MC_OK) return rc;
        }
    }
    trc.lap(1);
    if (trace_dev) cudaEventRecord(c->trace_ev[3], st);
    if (wait_upload) MC_CUDA(c, cudaStreamWaitEvent(st, c->ev_up, 0));
    if (trace_dev) cudaEventRecord(c->trace_ev[4], st);
    // A pipelined call of a decomposed run ends after its drift: the flags of that drift are agreed upon at the NEXT call, where
    // they ride with the all-gather of the external forces (same NCCL launch) instead of costing an all-reduce per call.
    bool flags_arrive = false;
    int n_ranks_f = 1;
    if (gather_chunk) {
        if (defer && c->flags_ride) {
            int rank_f = 0;
            comm_rank_size(c, &rank_f, &n_ranks_f);
            if (!c->h_flags_all) MC_CUDA(c, cudaMallocHost(&c->h_flags_all, sizeof(int) * 2 * 1024));
            MC_REQUIRE(c, n_ranks_f <= 1024, "mc_step: more than 1024 ranks");
            if ((rc = comm_allgather_ext_and_flags(c, const_cast<float *>(d_ext), gather_chunk, c->rebuild_flag.p, c->h_flags_all)) != MC_OK) return rc;
            flags_arrive = true;
        } else if ((rc = comm_allgather_f32_inplace(c, const_cast<float *>(d_ext), gather_chunk)) != MC_OK) {
            return rc;
        }
    }
    trc.lap(2);
    if (trace_dev) cudaEventRecord(c->trace_ev[5], st);
    const bool fused = fused_ok && (brute || !c->list_compact);
    if (fused) {
        int *h_out = reinterpret_cast<int *>(c->h_pinned) + 16;
        if (!c->ev_step_a) {
            MC_CUDA(c, cudaEventCreate(&c->ev_step_a)); MC_CUDA(c, cudaEventCreate(&c->ev_step_b));
            MC_CUDA(c, cudaEventCreateWithFlags(&c->ev_flag[0], cudaEventDisableTiming));
            MC_CUDA(c, cudaEventCreateWithFlags(&c->ev_flag[1], cudaEventDisableTiming));
        }
        MC_CUDA(c, cudaEventRecord(c->ev_step_a, st));
        int remaining = n_steps;
        bool first_half = true;
        while (remaining > 0) {
            // (a rebuild inside this loop may have switched to compact rows: the kernel reads global-slot rows)
            if (!brute && c->list_compact && (rc = engine_ensure_list32(c)) != MC_OK) return rc;
            FusedArgs A{};
            A.n = (int)c->n;
            A.brute = brute ? 1 : 0;
            A.lanes = c->fused_lanes;
            if (brute) {
                const uint32_t stride = ((uint32_t)c->n + 7u) & ~7u;
                MC_CUDA(c, c->bl_list.ensure((size_t)c->n * stride));
                MC_CUDA(c, c->bl_count.ensure((size_t)c->n));
                MC_CUDA(c, c->bl_xref.ensure((size_t)c->n));
                MC_CUDA(c, c->bl_flags.ensure(4));
                MC_CUDA(c, cudaMemsetAsync(c->bl_flags.p, 0, 4 * sizeof(int), st));
                A.bl_list = c->bl_list.p; A.bl_count = c->bl_count.p; A.bl_stride = stride; A.bl_xref = c->bl_xref.p;
                A.bl_flags = reinterpret_cast<int *>(c->bl_flags.p);
                A.need_forces = c->forces_valid ? 0 : 1;
                A.bl_keep = (c->bl_valid && c->bl_n == c->n && c->bl_rl2 == list_radius(c) * list_radius(c)) ? 1 : 0;
                A.rl2 = list_radius(c) * list_radius(c);
                A.excl_start = c->have_excl ? c->excl_start.p : nullptr;
                A.excl_idx = c->have_excl ? c->excl_idx.p : nullptr;
            }
            A.xyzq = c->xyzq[c->cur].p; A.vel = c->vel[c->cur].p; A.force = c->force.p;
            A.xref = (!brute || c->list_valid) ? c->xref.p : nullptr;  // (brute: only to tell whether the host's list went stale)
            A.type = c->type[c->cur].p; A.flags = c->flags[c->cur].p; A.orig = c->orig[c->cur].p; A.slot_of_orig = c->slot_of_orig.p;
            A.nbr_start = c->nbr_start.p; A.nbr_count = c->nbr_count.p; A.nbr_list = c->nbr_list.p;
            A.ljtab = c->ljtab.p; A.p = make_params(c); A.lj_on = c->lj_disabled ? 0 : 1;
            A.p14_start = c->have_p14 ? c->p14_start.p : nullptr; A.p14_idx = c->have_p14 ? c->p14_idx.p : nullptr;
            A.s14_lj = c->scale14_lj; A.s14_q = c->scale14_q;
            A.bt.n_bonds = c->n_bonds; A.bt.n_angles = c->n_angles; A.bt.n_dihedrals = c->n_dihedrals;
            A.bt.bonds = c->bonds.p; A.bt.bond_kr0 = c->bond_kr0.p; A.bt.angles = c->angles.p; A.bt.angle_kt0 = c->angle_kt0.p;
            A.bt.dihedrals = c->dihedrals.p; A.bt.dihedral_prm = c->dihedral_prm.p;
            A.ext_force = d_ext; A.dt = dt; A.max_disp = 0.5f * c->skin; A.n_steps = remaining; A.first_half = first_half ? 1 : 0;
            A.rebuild_flag = c->rebuild_flag.p; A.out = h_out;
            const int coul = c->coul_disabled ? MC_COULOMB_NONE : c->coul_mode;
            h_out[0] = -1; h_out[1] = 0; h_out[2] = 0; h_out[3] = 0;
            static const bool fused_times = getenv("MC_FUSED_TIMES") != nullptr;
            if (fused_times) {  // debugging aid: phase time stamps of the launch, printed after it
                MC_CUDA(c, c->fused_dbg.ensure(64));
                MC_CUDA(c, cudaMemsetAsync(c->fused_dbg.p, 0, 64 * sizeof(unsigned long long), st));
                A.dbg = c->fused_dbg.p;
            }
            MC_CUDA(c, launch_md_fused(A, c->n_types > 1, coul, c->periodic, c->n_sms, st, &c->launches));
            MC_CUDA(c, cudaStreamSynchronize(st));
            const int done = h_out[0], fl = h_out[1];
            if (fused_times) {
                unsigned long long h[64];
                cudaMemcpy(h, c->fused_dbg.p, sizeof(h), cudaMemcpyDeviceToHost);
                fprintf(stderr, "[fused times, ns since launch start]");
                for (int k = 1; k < 64 && h[k]; ++k) fprintf(stderr, " %llu", h[k] - h[0]);
                fprintf(stderr, "\n");
            }
            if (done < 0) return fail(c, MC_E_CUDA, "mc_step: the fused step kernel did not report back");
            if (fl & 2) return fail(c, MC_E_INVALID, "mc_step: non-finite coordinates (the simulation blew up)");
            c->n_steps += done;
            c->steps_since_build += done;
            remaining -= done;
            if (brute) {
                // every step taken on the kernel's own list; the host's list only goes stale (rebuilt when somebody needs it)
                c->n_rebuilds += h_out[3];
                if (h_out[2]) c->list_valid = false;
                c->forces_valid = true;
                c->bl_valid = true;  // (cleared by everything else that moves, reorders or redefines atoms: MC_FLUSH, the builds, the other step paths)
                c->bl_n = c->n;
                c->bl_rl2 = list_radius(c) * list_radius(c);
                break;
            }
            if (!(fl & 1)) break;  // all steps taken, closing half kick applied in the kernel; forces belong to the positions
            // the last drift tripped the displacement criterion: its forces are still those of the previous positions
            c->forces_valid = false;
            if ((rc = engine_build_list(c)) != MC_OK) return rc;
            if ((rc = engine_launch_forces(c, false)) != MC_OK) return rc;
            first_half = false;
            if (remaining == 0) {
                launch_kick_drift((int)c->n, c->xyzq[c->cur].p, c->vel[c->cur].p, c->force.p, d_ext, c->orig[c->cur].p, c->flags[c->cur].p,
                                  c->xref.p, 0.5f * dt, 0.f, 0.f, 0.f, c->rebuild_flag.p, st, &c->launches);
            }
        }
        c->forces_have_energy = false;
        MC_CUDA(c, cudaEventRecord(c->ev_step_b, st));
        MC_CUDA(c, cudaStreamSynchronize(st));
        float ms_f = 0.f;
        MC_CUDA(c, cudaEventElapsedTime(&ms_f, c->ev_step_a, c->ev_step_b));
        c->last_step_ms = ms_f;
        return MC_OK;
    }
    // Velocity Verlet, two kernels per step: [kick + drift] and [pair forces].  The second half
    // kick of step s and the first half kick of step s+1 are one full kick in the same launch.
    // The rebuild decision is pipelined: kick_drift raises the flag with a look-ahead margin, the
    // flag travels to pinned host memory asynchronously, and the host acts on the flag of the
    // PREVIOUS step while the GPU is already busy -- no per-step stream synchronisation.
    const float max_disp = 0.5f * c->skin;
    // the host acts on the flag of step s-1 before the pair kernel of step s: the stale list is last used one
    // drift after the flag could first have been raised, so one drift of margin is needed; 1.5 leaves room
    // for the change of velocity within that step
    const float lookahead = pipelined ? 1.5f : 0.f;
    int *h_flag = reinterpret_cast<int *>(c->h_pinned) + 8;  // two slots
    if (!c->ev_step_a) {
        MC_CUDA(c, cudaEventCreate(&c->ev_step_a)); MC_CUDA(c, cudaEventCreate(&c->ev_step_b));
        MC_CUDA(c, cudaEventCreateWithFlags(&c->ev_flag[0], cudaEventDisableTiming));
        MC_CUDA(c, cudaEventCreateWithFlags(&c->ev_flag[1], cudaEventDisableTiming));
    }
    MC_CUDA(c, cudaEventRecord(c->ev_step_a, st));
    bool have_prev = false, skip_prev = false;
    int tag[2] = {0, 0};
    const bool fused_halo = c->comm_active && comm_peer_direct(c) && c->halo_fused;
    for (int s = 0; s < n_steps; ++s) {
        HaloSplit split{};
        c->prof_now = c->prof_every <= 1 || (s % c->prof_every) == 0;
        {
            TimedRegion tr(c, c->integ_acc, true);
            const size_t r0 = (size_t)c->row0;
            if (fused_halo) {
                HaloPush push{};
                comm_step_descriptors(c, c->steps_since_build + 1 >= comm_interval(c), &push, &split);
                launch_kick_drift_halo((int)c->n_rows_sorted(), c->xyzq[c->cur].p + r0, c->vel[c->cur].p + r0, c->force.p + r0,
                                       d_ext, c->orig[c->cur].p + r0, c->flags[c->cur].p + r0, c->xref.p + r0,
                                       (s == 0 || constrained) ? 0.5f * dt : dt, dt, max_disp, c->rebuild_flag.p, push, st, &c->launches);
            } else {
                if (pipelined) tag[s & 1] = (++c->flag_tag) & 0x1fffffff;
                launch_kick_drift((int)c->n_rows_sorted(), c->xyzq[c->cur].p + r0, c->vel[c->cur].p + r0, c->force.p + r0, d_ext,
                                  c->orig[c->cur].p + r0, c->flags[c->cur].p + r0, c->xref.p + r0, (s == 0 || constrained) ? 0.5f * dt : dt, dt,
                                  max_disp, lookahead, c->rebuild_flag.p, st, &c->launches,
                                  pipelined ? reinterpret_cast<uint32_t *>(c->rebuild_flag.p + 2) : nullptr,
                                  pipelined ? h_flag + (s & 1) : nullptr, tag[s & 1]);
            }
            tr.stop();
        }
        if (c->n_waters > 0 || c->n_hclusters > 0) {
            // virial of this step's constraint forces (mc_get_pressure): one fp64 atomic per warp inside the kernels
            MC_CUDA(c, c->cons_virial.ensure(1));
            MC_CUDA(c, cudaMemsetAsync(c->cons_virial.p, 0, sizeof(double), st));
            c->cons_virial_valid = true;
        }
        if (c->n_waters > 0)
            launch_settle(c->n_waters, c->waters.p, c->slot_of_orig.p, c->xyzq[c->cur].p, c->vel[c->cur].p, c->water_m_o,
                          c->water_m_h, c->water_d_oh, c->water_d_hh, make_params(c), dt, c->cons_virial.p, st, &c->launches);
        if (c->n_hclusters > 0)
            launch_shake_h(c->n_hclusters, c->hclusters.p, c->hdist.p, c->slot_of_orig.p, c->xyzq[c->cur].p, c->vel[c->cur].p,
                           make_params(c), dt, c->shake_tol, c->shake_fail.p, c->cons_virial.p, st, &c->launches);
        if (c->n_vsites > 0)
            launch_vsite_construct(c->n_vsites, c->vsites.p, c->slot_of_orig.p, c->xyzq[c->cur].p, c->vsite_a, c->vsite_b,
                                   make_params(c), st, &c->launches);
        if (c->com_every > 0 && !c->comm_active && (c->n_steps + 1) % c->com_every == 0) {
            MC_CUDA(c, c->com_partial.ensure((size_t)com_partial_elems()));
            const size_t r0 = (size_t)c->row0;
            launch_remove_com((int)c->n_rows_sorted(), c->vel[c->cur].p + r0, c->flags[c->cur].p + r0, c->com_partial.p, st, &c->launches);
        }
        if (c->langevin) {
            const float c1 = std::exp(-c->lgv_gamma * dt);
            const size_t r0 = (size_t)c->row0;
            launch_langevin_ou((int)c->n_rows_sorted(), c->vel[c->cur].p + r0, c->orig[c->cur].p + r0, c->flags[c->cur].p + r0, c1,
                               std::sqrt(std::max(0.f, 1.f - c1 * c1)), (float)MC_KB * c->lgv_temperature, c->lgv_seed, c->lgv_step++,
                               st, &c->launches);
        }
        if (c->csvr) {
            // kinetic energy of the half-step velocities (two small reduction launches), then lambda, then the scaling
            MC_CUDA(c, c->red_partial.ensure((size_t)energy_partial_elems()));
            MC_CUDA(c, c->red_out.ensure(4));
            MC_CUDA(c, c->csvr_lambda.ensure(1));
            const size_t r0 = (size_t)c->row0;
            launch_energy_reduce((int)c->n_rows_sorted(), c->force.p + r0, c->vel[c->cur].p + r0, c->flags[c->cur].p + r0, c->red_partial.p, c->red_out.p, st,
                                 &c->launches);
            // decomposed: every rank reduced its owned atoms; one 24-byte all-reduce on the stream gives all of them the same
            // kinetic energy and mobile-atom count, hence the same scaling factor (a function of seed, step and that energy)
            if (c->comm_active && (rc = comm_allreduce_dev_f64(c, c->red_out.p, 3)) != MC_OK) return rc;
            launch_csvr((int)c->n_rows_sorted(), c->vel[c->cur].p + r0, c->red_out.p, MC_KB * (double)c->lgv_temperature,
                        std::exp(-(double)c->lgv_gamma * (double)dt), 3.0 * (double)c->n_waters + (double)c->n_hconstraints, c->lgv_seed, c->lgv_step++,
                        c->csvr_lambda.p, st, &c->launches);
        }
        c->steps_since_build++;
        if (defer && s == n_steps - 1) {
            // last step of a pipelined call: stop after the drift.  Its flag word is read after the final
            // synchronisation below (no rebuild has happened since it was written, so it is never stale).
            if (!c->ev_drift) MC_CUDA(c, cudaEventCreateWithFlags(&c->ev_drift, cudaEventDisableTiming));
            MC_CUDA(c, cudaEventRecord(c->ev_drift, st));  // positions final: a snapshot need not queue behind the early force evaluation
            c->drift_event_valid = true;
            c->tail_rebuild = c->comm_active && c->steps_since_build >= comm_interval(c);
            c->tail_use_split = fused_halo && !c->tail_rebuild;
            c->tail_split = split;
            have_prev = true;
            skip_prev = false;
            c->forces_valid = false;
            c->n_steps++;
            break;
        }
        bool rebuild = false;
        if (c->comm_active) {
            rebuild = c->steps_since_build >= comm_interval(c);
        } else if (c->rebuild_every > 0) {
            rebuild = c->steps_since_build >= c->rebuild_every;
        } else if (pipelined) {
            // kick_drift publishes {tag, flag} into pinned host memory itself; nothing is queued between it and the
            // pair kernel.  The word of the PREVIOUS step was written one pair kernel ago.
            if (have_prev && !skip_prev) {
                if ((rc = wait_flag_tag(c, h_flag + ((s - 1) & 1), tag[(s - 1) & 1])) != MC_OK) return rc;
                rebuild = (h_flag[(s - 1) & 1] & 1) != 0;
                if (h_flag[(s - 1) & 1] & 2) return fail(c, MC_E_INVALID, "mc_step: non-finite coordinates (the simulation blew up)");
            }
            have_prev = true;
            skip_prev = rebuild;  // the flag copied just above still refers to the old reference positions
        } else {
            MC_CUDA(c, cudaMemcpyAsync(h_flag, c->rebuild_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            MC_CUDA(c, cudaStreamSynchronize(st));
            rebuild = (*h_flag & 1) != 0;
            if (*h_flag & 2) return fail(c, MC_E_INVALID, "mc_step: non-finite coordinates (the simulation blew up)");
            if (c->comm_active && (rc = comm_agree_flag(c, &rebuild)) != MC_OK) return rc;
        }
        if (rebuild) {
            rc = c->comm_active ? comm_rebuild(c) : engine_build_list(c);
            if (rc != MC_OK) return rc;
        } else if (c->comm_active && !fused_halo && (rc = comm_halo_positions(c)) != MC_OK) {
            return rc;
        }
        const bool baro_now = baro && (c->n_steps + 1) % c->baro_every == 0;
        if ((rc = engine_launch_forces(c, baro_now, fused_halo && !rebuild ? &split : nullptr)) != MC_OK) return rc;
        if (baro_now) {
            // pressure of the positions just reached -> scale box and coordinates -> rebuild -> forces of the scaled system
            if ((rc = apply_barostat(c, dt)) != MC_OK) return rc;
            skip_prev = true;  // the flag word of this step refers to the reference positions of the old list
        }
        if (constrained) {
            // closing half kick of THIS step, then RATTLE's velocity stage on the positions of this step
            const size_t r0 = (size_t)c->row0;
            launch_kick_drift((int)c->n_rows_sorted(), c->xyzq[c->cur].p + r0, c->vel[c->cur].p + r0, c->force.p + r0, d_ext,
                              c->orig[c->cur].p + r0, c->flags[c->cur].p + r0, c->xref.p + r0, 0.5f * dt, 0.f, 0.f, 0.f,
                              c->rebuild_flag.p, st, &c->launches);
            launch_rattle_velocities(c->n_waters, c->waters.p, c->n_hclusters, c->hclusters.p, c->slot_of_orig.p, c->xyzq[c->cur].p,
                                     c->vel[c->cur].p, make_params(c), st, &c->launches);
        }
        c->n_steps++;
    }
    if (defer) {
        c->tail_pending = true;
        c->tail_dt = dt;
        c->tail_ext = d_ext;
    } else if (n_steps > 0 && !constrained) {
        TimedRegion tr(c, c->integ_acc);
        const size_t r0 = (size_t)c->row0;
        launch_kick_drift((int)c->n_rows_sorted(), c->xyzq[c->cur].p + r0, c->vel[c->cur].p + r0, c->force.p + r0, d_ext,
                          c->orig[c->cur].p + r0, c->flags[c->cur].p + r0, c->xref.p + r0, 0.5f * dt, 0.f, 0.f, 0.f,
                          c->rebuild_flag.p, st, &c->launches);
        tr.stop();
    }
    c->prof_now = true;
    MC_CUDA(c, cudaEventRecord(c->ev_step_b, st));
    trc.lap(3);
    // Fixed schedules (rebuild_every > 0, decomposed runs): the displacement flag is the safety net.  It rides behind the
    // last kernel -- on a decomposed handle as the maximum over all ranks (one 8-byte all-reduce per CALL, not per step),
    // so that every rank takes the same decision -- and is read after the one synchronisation of this call.
    int *h_agree = reinterpret_cast<int *>(c->h_pinned) + 12;
    // (a pipelined decomposed call with external forces leaves its flags to the next call's all-gather, see above)
    const bool ride = defer && c->comm_active && gather_chunk != 0;
    c->flags_ride = ride;
    const bool check_flag = (c->rebuild_every > 0 || c->comm_active) && n_steps > 0 && !ride;
    if (check_flag) {
        if (c->comm_active) {
            if ((rc = comm_reduce_flags_async(c, c->rebuild_flag.p, h_agree)) != MC_OK) return rc;
        } else {
            MC_CUDA(c, cudaMemcpyAsync(h_agree, c->rebuild_flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        }
    }
    MC_CUDA(c, cudaGetLastError());
    trc.lap(4);
    StepEpilogue E;
    E.pipelined = pipelined; E.n_steps = n_steps; E.skip_prev = skip_prev; E.check_flag = check_flag; E.flags_arrive = flags_arrive;
    E.n_ranks_f = n_ranks_f; E.trace_dev = trace_dev && defer;
    if (defer && c->lazy_sync) {
        // A pipelined call has nothing to hand back but positions, and those are read in stream order (snapshots, mc_get_positions):
        // it returns as soon as the caller's array is free again (UploadGuard) and leaves its kernels running.  What the
        // synchronisation used to deliver -- the flag words, the timing -- is collected when the next call (or whoever closes the
        // open step) comes back: the host side of a per-step loop then runs under the kernels instead of between them.
        E.pending = true;
        c->epi = E;
        if (c->early_tail) {
            const int64_t builds = c->n_rebuilds;
            if ((rc = tail_forces(c)) != MC_OK) return rc;
            // a rebuild here comes after the drift whose flag word is still to be read: that word refers to the old reference positions
            if (c->n_rebuilds != builds) c->epi.skip_prev = true;
        }
        return lazy_rc;
    }
    const int rc_e = step_epilogue(c, E);
    trc.lap(5);
    if (defer && c->early_tail && rc_e >= 0 && (rc = tail_forces(c)) != MC_OK) return rc;
    return rc_e != MC_OK ? rc_e : lazy_rc;
}

extern "C" double mc_last_step_ms(mc_ctx *c) { return c ? c->last_step_ms : 0.0; }

// ---- energy minimisation -----------------------------------------------------------------------------------
// md.minimize_energy(dev, max_iters, None) (reference ui/mol_editor.rs:375, mol_alignment.rs:356,
// properties/sol_shrinking_box.rs:962): steepest descent with an adaptive step, built from the step path's own
// kernels -- velocities are zeroed, one kick + drift of length tau moves every atom along F/m (a quenched MD step:
// dx = 418.4 F/m tau^2), the displacement flag triggers list rebuilds exactly as in mc_step, and the move is kept
// when the potential energy went down (tau x 1.2) or undone from a copy in original order (tau x 0.5).
// STATUS: host logic written after round 1's GPU budget was spent; not yet run on hardware.
extern "C" int mc_minimize_energy(mc_ctx *c, int max_iters, int *iters_accepted, double *e_initial, double *e_final) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, max_iters >= 0, "mc_minimize_energy: max_iters >= 0 required");
    MC_REQUIRE(c, !c->comm_active, "mc_minimize_energy: not available on a decomposed handle yet");
    MC_REQUIRE(c, c->n_waters == 0 && c->n_hclusters == 0, "mc_minimize_energy: constraints are not applied by the minimiser; clear them first");
    int rc = mc_compute_forces(c);
    if (rc != MC_OK) return rc;
    mc_energy en;
    if ((rc = mc_get_energy(c, &en)) != MC_OK) return rc;
    double e_cur = en.energy_potential;
    if (e_initial) *e_initial = e_cur;
    const int n = (int)c->n;
    cudaStream_t st = c->st;
    MC_CUDA(c, c->min_x.ensure((size_t)std::max(n, 1)));
    MC_CUDA(c, c->min_v.ensure((size_t)std::max(n, 1)));
    launch_gather_to_orig(n, c->vel[c->cur].p, c->orig[c->cur].p, c->min_v.p, st, &c->launches);
    float tau = 0.001f;  // ps
    int accepted = 0, small = 0;
    int *h_flag = reinterpret_cast<int *>(c->h_pinned) + 8;
    for (int it = 0; it < max_iters && n > 0; ++it) {
        launch_gather_to_orig(n, c->xyzq[c->cur].p, c->orig[c->cur].p, c->min_x.p, st, &c->launches);
        launch_zero_velocities(n, c->vel[c->cur].p, st, &c->launches);
        launch_kick_drift(n, c->xyzq[c->cur].p, c->vel[c->cur].p, c->force.p, nullptr, c->orig[c->cur].p, c->flags[c->cur].p, c->xref.p,
                          tau, tau, 0.5f * c->skin, 0.f, c->rebuild_flag.p, st, &c->launches);
        // virtual sites (massless, static: the drift leaves them where they were) follow their parents, as after every drift of mc_step
        if (c->n_vsites > 0)
            launch_vsite_construct(c->n_vsites, c->vsites.p, c->slot_of_orig.p, c->xyzq[c->cur].p, c->vsite_a, c->vsite_b,
                                   make_params(c), st, &c->launches);
        MC_CUDA(c, cudaMemcpyAsync(h_flag, c->rebuild_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        MC_CUDA(c, cudaStreamSynchronize(st));
        const bool blew_up = (*h_flag & 2) != 0;
        if (*h_flag & 1) c->list_valid = false;
        c->forces_valid = false;
        double e_try = 0.0;
        if (!blew_up) {
            if ((rc = mc_compute_forces(c)) != MC_OK) return rc;
            if ((rc = mc_get_energy(c, &en)) != MC_OK) return rc;
            e_try = en.energy_potential;
        }
        if (!blew_up && e_try <= e_cur) {
            small = (e_cur - e_try) <= 1e-9 * std::max(1.0, std::fabs(e_cur)) ? small + 1 : 0;
            e_cur = e_try;
            ++accepted;
            tau = std::min(tau * 1.2f, 0.02f);
            if (small >= 3) break;  // converged: three accepted moves in a row changed nothing
        } else {
            // undo: positions back from the copy in original order (the list may have been rebuilt meanwhile, so
            // slots are looked up afresh), list rebuilt, forces of the restored positions re-evaluated
            launch_scatter_from_orig(n, c->min_x.p, c->orig[c->cur].p, c->xyzq[c->cur].p, 0, st, &c->launches);
            MC_CUDA(c, cudaMemsetAsync(c->rebuild_flag.p, 0, sizeof(int), st));
            c->list_valid = false;
            c->forces_valid = false;
            if ((rc = mc_compute_forces(c)) != MC_OK) return rc;
            tau *= 0.5f;
            if (tau < 1e-7f) break;
        }
    }
    launch_scatter_from_orig(n, c->min_v.p, c->orig[c->cur].p, c->vel[c->cur].p, 0, st, &c->launches);
    MC_CUDA(c, cudaStreamSynchronize(st));
    if (iters_accepted) *iters_accepted = accepted;
    if (e_final) *e_final = e_cur;
    return MC_OK;
}

// ---- read-back -------------------------------------------------------------------------------------------

static int read_sorted_to_orig(mc_ctx *c, const float4 *sorted, mc_float4 *out) {
    const int64_t n = c->n_rows_sorted();
    if (n == 0) return MC_OK;
    MC_CUDA(c, c->stage.ensure((size_t)c->n_global));
    if (c->comm_active) MC_CUDA(c, cudaMemsetAsync(c->stage.p, 0, sizeof(float4) * c->n_global, c->st));
    launch_gather_to_orig((int)n, sorted + c->row0, c->orig[c->cur].p + c->row0, c->stage.p, c->st, &c->launches);
    if (c->comm_active) { int rc = comm_allreduce_f4(c, c->stage.p, c->n_global); if (rc != MC_OK) return rc; }
    MC_CUDA(c, cudaMemcpyAsync(out, c->stage.p, sizeof(float4) * c->n_global, cudaMemcpyDeviceToHost, c->st));
    MC_CUDA(c, cudaStreamSynchronize(c->st));
    return MC_OK;
}

extern "C" int mc_get_positions(mc_ctx *c, mc_float4 *out) {
    if (!c || !out) return MC_E_INVALID;
    cudaSetDevice(c->device);
    return read_sorted_to_orig(c, c->xyzq[c->cur].p, out);
}

extern "C" int mc_get_velocities(mc_ctx *c, mc_float4 *out) {
    if (!c || !out) return MC_E_INVALID;
    MC_FLUSH_OBS(c);
    cudaSetDevice(c->device);
    return read_sorted_to_orig(c, c->vel[c->cur].p, out);
}

extern "C" int mc_get_forces(mc_ctx *c, mc_float4 *out) {
    if (!c || !out) return MC_E_INVALID;
    MC_FLUSH_OBS(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, c->forces_valid, "mc_get_forces: no force evaluation since the last change; call mc_compute_forces");
    return read_sorted_to_orig(c, c->force.p, out);
}

// ---- asynchronous snapshot hand-off --------------------------------------------------
// The reference queues Snapshot{atom_posits, ...} objects while the integrator keeps going
// (src/md/mod.rs:118-152 flush_snapshot_queues, md/trajectory.rs:160-204).  Here: the positions are
// copied into one of two device staging buffers on the compute stream (a gather into original order on
// a single GPU; the owned block + its original ids on a decomposed rank), and a second stream moves
// the staging buffer to the caller's host buffer while the next steps already run.

static int snapshot_begin_impl(mc_ctx *c, mc_float4 *out_positions, mc_float4 *out_velocities, int32_t *out_ids, int64_t *n_out) {
    if (!c || !out_positions) return MC_E_INVALID;
    if (out_velocities) MC_FLUSH_OBS(c);  // velocities are only final once the step mc_step may have left open is closed
    cudaSetDevice(c->device);
    MC_REQUIRE(c, !c->comm_active || out_ids, "mc_snapshot_begin: a decomposed handle returns its owned atoms and needs out_ids");
    if (!c->st_copy) {
        MC_CUDA(c, cudaStreamCreateWithFlags(&c->st_copy, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            MC_CUDA(c, cudaEventCreateWithFlags(&c->ev_snap_staged[b], cudaEventDisableTiming));
            MC_CUDA(c, cudaEventCreateWithFlags(&c->ev_snap_done[b], cudaEventDisableTiming));
        }
    }
    const int k = c->snap_k;
    c->snap_k ^= 1;
    const int64_t rows = c->n_rows_sorted();
    const int64_t n = c->comm_active ? rows : c->n_global;
    if (n_out) *n_out = n;
    if (n == 0) return MC_OK;
    // the copy that last used this staging buffer must have drained before it is overwritten
    if (c->snap_pending[k]) MC_CUDA(c, cudaStreamWaitEvent(c->st, c->ev_snap_done[k], 0));
    MC_CUDA(c, c->snap_stage[k].ensure((size_t)n));
    if (c->comm_active) {
        MC_CUDA(c, c->snap_ids[k].ensure((size_t)n));
        MC_CUDA(c, cudaMemcpyAsync(c->snap_stage[k].p, c->xyzq[c->cur].p + c->row0, sizeof(float4) * n, cudaMemcpyDeviceToDevice, c->st));
        MC_CUDA(c, cudaMemcpyAsync(c->snap_ids[k].p, c->orig[c->cur].p + c->row0, sizeof(int) * n, cudaMemcpyDeviceToDevice, c->st));
    } else {
        launch_gather_to_orig((int)rows, c->xyzq[c->cur].p, c->orig[c->cur].p, c->snap_stage[k].p, c->st, &c->launches);
    }
    if (out_velocities) {
        MC_CUDA(c, c->snap_stage_v[k].ensure((size_t)n));
        if (c->comm_active)
            MC_CUDA(c, cudaMemcpyAsync(c->snap_stage_v[k].p, c->vel[c->cur].p + c->row0, sizeof(float4) * n, cudaMemcpyDeviceToDevice, c->st));
        else
            launch_gather_to_orig((int)rows, c->vel[c->cur].p, c->orig[c->cur].p, c->snap_stage_v[k].p, c->st, &c->launches);
    }
    MC_CUDA(c, cudaEventRecord(c->ev_snap_staged[k], c->st));
    MC_CUDA(c, cudaStreamWaitEvent(c->st_copy, c->ev_snap_staged[k], 0));
    MC_CUDA(c, cudaMemcpyAsync(out_positions, c->snap_stage[k].p, sizeof(float4) * n, cudaMemcpyDeviceToHost, c->st_copy));
    if (out_velocities)
        MC_CUDA(c, cudaMemcpyAsync(out_velocities, c->snap_stage_v[k].p, sizeof(float4) * n, cudaMemcpyDeviceToHost, c->st_copy));
    if (c->comm_active) MC_CUDA(c, cudaMemcpyAsync(out_ids, c->snap_ids[k].p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->st_copy));
    MC_CUDA(c, cudaEventRecord(c->ev_snap_done[k], c->st_copy));
    c->snap_pending[k] = true;
    return MC_OK;
}

// Snapshot.atom_posits is a Vec<Vec3F32> (reference src/md/trajectory.rs:160-204): 12 bytes per atom are what the viewer
// needs per frame -- a quarter less PCIe traffic than the float4 record; and the ids of a decomposed rank only change at a
// list rebuild, so they travel only when the caller has not yet seen this layout (layout_epoch).
extern "C" int mc_snapshot_begin_xyz(mc_ctx *c, float *out_xyz, int32_t *out_ids, int64_t *n_out, int64_t *layout_epoch) {
    if (!c || !out_xyz) return MC_E_INVALID;
    cudaSetDevice(c->device);
    if (!c->st_copy) {
        MC_CUDA(c, cudaStreamCreateWithFlags(&c->st_copy, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            MC_CUDA(c, cudaEventCreateWithFlags(&c->ev_snap_staged[b], cudaEventDisableTiming));
            MC_CUDA(c, cudaEventCreateWithFlags(&c->ev_snap_done[b], cudaEventDisableTiming));
        }
    }
    const int k = c->snap_k;
    c->snap_k ^= 1;
    const int64_t rows = c->n_rows_sorted();
    const int64_t n = c->comm_active ? rows : c->n_global;
    if (n_out) *n_out = n;
    // *layout_epoch on entry: the layout the caller already holds the ids of (they travel only when it is another one)
    const bool want_ids = out_ids != nullptr && (!layout_epoch || *layout_epoch != c->n_rebuilds);
    if (layout_epoch) *layout_epoch = c->n_rebuilds;
    if (n == 0) return MC_OK;
    // A pipelined mc_step that has already launched the open step's force evaluation (early_tail) recorded an event right after
    // its drift: the staging then runs on the copy stream behind that event instead of queueing behind the force kernel, and
    // the engine stream waits for the staging before anything may move the atoms again.
    const bool side = c->tail_pending && c->drift_event_valid && c->forces_valid && !(c->comm_active && c->tail_rebuild);
    cudaStream_t ss = side ? c->st_copy : c->st;
    if (side) MC_CUDA(c, cudaStreamWaitEvent(c->st_copy, c->ev_drift, 0));
    else if (c->snap_pending[k]) MC_CUDA(c, cudaStreamWaitEvent(c->st, c->ev_snap_done[k], 0));
    // (head-room: a decomposed rank owns a few atoms more or fewer after every rebuild, and growing a buffer means cudaFree)
    MC_CUDA(c, c->snap_stage[k].ensure((size_t)n + (c->comm_active ? (size_t)n / 8 + 1024 : 0)));  // float4 elements: 3n floats fit
    float *stage = reinterpret_cast<float *>(c->snap_stage[k].p);
    launch_pack_xyz((int)rows, c->xyzq[c->cur].p + c->row0, c->comm_active ? nullptr : c->orig[c->cur].p + c->row0, stage, ss, &c->launches);
    if (c->comm_active && want_ids) {
        MC_CUDA(c, c->snap_ids[k].ensure((size_t)n + (size_t)n / 8 + 1024));
        MC_CUDA(c, cudaMemcpyAsync(c->snap_ids[k].p, c->orig[c->cur].p + c->row0, sizeof(int) * n, cudaMemcpyDeviceToDevice, ss));
    }
    MC_CUDA(c, cudaEventRecord(c->ev_snap_staged[k], ss));
    if (side) MC_CUDA(c, cudaStreamWaitEvent(c->st, c->ev_snap_staged[k], 0));
    else MC_CUDA(c, cudaStreamWaitEvent(c->st_copy, c->ev_snap_staged[k], 0));
    MC_CUDA(c, cudaMemcpyAsync(out_xyz, stage, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->st_copy));
    if (c->comm_active && want_ids) MC_CUDA(c, cudaMemcpyAsync(out_ids, c->snap_ids[k].p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->st_copy));
    MC_CUDA(c, cudaEventRecord(c->ev_snap_done[k], c->st_copy));
    c->snap_pending[k] = true;
    return MC_OK;
}

extern "C" int mc_snapshot_begin(mc_ctx *c, mc_float4 *out_positions, int32_t *out_ids, int64_t *n_out) {
    return snapshot_begin_impl(c, out_positions, nullptr, out_ids, n_out);
}

// Snapshot{atom_posits, atom_velocities, ..} (reference src/md/trajectory.rs:160-204): positions AND velocities
extern "C" int mc_snapshot_begin_pv(mc_ctx *c, mc_float4 *out_positions, mc_float4 *out_velocities, int32_t *out_ids, int64_t *n_out) {
    if (!out_velocities) return MC_E_INVALID;
    return snapshot_begin_impl(c, out_positions, out_velocities, out_ids, n_out);
}

extern "C" int mc_snapshot_wait(mc_ctx *c) {
    if (!c) return MC_E_INVALID;
    cudaSetDevice(c->device);
    if (!c->st_copy) return MC_OK;
    // the oldest outstanding snapshot is the one in the buffer the next begin would use
    for (int t = 0; t < 2; ++t) {
        const int k = c->snap_k ^ t;
        if (c->snap_pending[k]) {
            MC_CUDA(c, cudaEventSynchronize(c->ev_snap_done[k]));
            c->snap_pending[k] = false;
            return MC_OK;
        }
    }
    return MC_OK;
}

extern "C" int mc_get_energy(mc_ctx *c, mc_energy *out) {
    if (!c || !out) return MC_E_INVALID;
    MC_FLUSH_OBS(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, c->forces_valid, "mc_get_energy: no force evaluation since the last change; call mc_compute_forces");
    if (!c->forces_have_energy) {
        // the step path skips the energy row sums; evaluate them now on the unchanged positions
        int rc = engine_launch_forces(c, true);
        if (rc != MC_OK) return rc;
    }
    MC_CUDA(c, c->red_partial.ensure((size_t)energy_partial_elems()));
    MC_CUDA(c, c->red_out.ensure(4));
    launch_energy_reduce((int)c->n_rows_sorted(), c->force.p + c->row0, c->vel[c->cur].p + c->row0, c->flags[c->cur].p + c->row0, c->red_partial.p, c->red_out.p, c->st,
                         &c->launches);
    double h[3];
    MC_CUDA(c, cudaMemcpyAsync(h, c->red_out.p, sizeof(h), cudaMemcpyDeviceToHost, c->st));
    MC_CUDA(c, cudaStreamSynchronize(c->st));
    if (c->comm_active) { int rc = comm_allreduce3(c, h); if (rc != MC_OK) return rc; }
    memset(out, 0, sizeof(*out));
    out->energy_potential_nonbonded = 0.5 * h[0];  // every pair sits in two rows
    out->energy_potential_bonded = 0.0;
    if (c->n_bonds + c->n_angles + c->n_dihedrals > 0) {
        double hb[3];
        MC_CUDA(c, cudaMemcpy(hb, c->bonded_e.p, sizeof(hb), cudaMemcpyDeviceToHost));
        if (c->comm_active) {
            // every rank added the share of each term that belongs to its owned atoms
            int rc = comm_allreduce3(c, hb);
            if (rc != MC_OK) return rc;
            int bad = 0;
            if (c->bonded_missing.p) MC_CUDA(c, cudaMemcpy(&bad, c->bonded_missing.p, sizeof(int), cudaMemcpyDeviceToHost));
            if (bad) {
                MC_CUDA(c, cudaMemset(c->bonded_missing.p, 0, sizeof(int)));
                return fail(c, MC_E_INVALID, "mc_get_energy: a bonded term reaches beyond this rank's ghost layer (bonded partners must lie within cutoff + skin)");
            }
        }
        out->energy_bond = hb[0]; out->energy_angle = hb[1]; out->energy_dihedral = hb[2];
        out->energy_potential_bonded = hb[0] + hb[1] + hb[2];
    }
    if (c->pme.planned && c->periodic && !c->comm_active && c->coul_mode == MC_COULOMB_ERFC && !c->coul_disabled) {
        double hp[2];
        MC_CUDA(c, cudaMemcpy(hp, c->pme.energy, sizeof(hp), cudaMemcpyDeviceToHost));
        // reciprocal sum + self term -alpha/sqrt(pi) sum q^2 + erf correction of the excluded pairs
        out->energy_pme = hp[0] - (double)c->alpha * 0.5641895835477563 * c->pme.self_q2 + (c->have_excl ? hp[1] : 0.0);
        out->energy_potential_nonbonded += out->energy_pme;
    }
    out->energy_potential = out->energy_potential_nonbonded + out->energy_potential_bonded;
    if (c->periodic) {
        out->volume = (double)c->ext[0] * (double)c->ext[1] * (double)c->ext[2];
        out->density = out->volume > 0.0 ? c->total_mass * 1.66053907 / out->volume : 0.0;  // amu/A^3 -> g/cm^3
    }
    out->energy_kinetic = h[1] / (double)MC_ACCEL_CONV;
    // each rigid water removes three degrees of freedom, each constrained bond one
    // ... and removing the centre-of-mass drift takes three more
    const double dof = 3.0 * h[2] - 3.0 * (double)c->n_waters - (double)c->n_hconstraints - (c->com_every > 0 ? 3.0 : 0.0);
    out->temperature = dof > 0 ? 2.0 * out->energy_kinetic / (dof * MC_KB) : 0.0;
    return MC_OK;
}

// SnapshotEnergyData.pressure (reference ui/panels/md_viewer.rs:202-256; the barostat's input, ui/panels/md.rs:517-556):
// P = (2 KE + W) / 3V with the virial W = sum r_ij . f_ij of the nonbonded pairs (one extra pass over the list, on
// demand only), the scaled 1-4 pairs, the bonded terms and, with SPME, the reciprocal sum and its excluded-pair
// correction (accumulated next to the energies of the last force evaluation).  bar = kcal/mol/A^3 x 69476.95.
// forces of the current positions must be valid and evaluated with energies (bonded / SPME virials sit next to them)
static int compute_pressure(mc_ctx *c, double *pressure_bar, double *virial) {
    const bool constrained = c->n_waters > 0 || c->n_hclusters > 0;
    MC_CUDA(c, c->red_partial.ensure((size_t)energy_partial_elems()));
    MC_CUDA(c, c->red_out.ensure(4));
    launch_energy_reduce((int)c->n_rows_sorted(), c->force.p + c->row0, c->vel[c->cur].p + c->row0, c->flags[c->cur].p + c->row0, c->red_partial.p, c->red_out.p, c->st,
                         &c->launches);
    const int coul = c->coul_disabled ? MC_COULOMB_NONE : c->coul_mode;
    { int rc32 = engine_ensure_list32(c); if (rc32 != MC_OK) return rc32; }
    launch_virial((int)c->n_rows_sorted(), (int)c->row0, c->xyzq[c->cur].p, c->type[c->cur].p, c->orig[c->cur].p, c->slot_of_orig.p,
                  c->nbr_start.p, c->nbr_count.p, c->nbr_list.p, c->have_p14 ? c->p14_start.p : nullptr, c->have_p14 ? c->p14_idx.p : nullptr,
                  c->ljtab.p, make_params(c), c->lj_disabled ? 0 : 1, coul, c->scale14_lj, c->scale14_q, c->red_out.p + 3, c->st,
                  &c->launches);
    double h[4];
    MC_CUDA(c, cudaMemcpyAsync(h, c->red_out.p, sizeof(h), cudaMemcpyDeviceToHost, c->st));
    MC_CUDA(c, cudaStreamSynchronize(c->st));
    double w = h[3];
    if (c->n_bonds + c->n_angles + c->n_dihedrals > 0) {
        double hb[4];
        MC_CUDA(c, cudaMemcpy(hb, c->bonded_e.p, sizeof(hb), cudaMemcpyDeviceToHost));
        w += hb[3];
    }
    if (c->pme.planned && coul == MC_COULOMB_ERFC) {
        double hp[4];
        MC_CUDA(c, cudaMemcpy(hp, c->pme.energy, sizeof(hp), cudaMemcpyDeviceToHost));
        w += hp[2] + (c->have_excl ? hp[3] : 0.0);  // the self term does not depend on the volume
    }
    if (c->comm_active) {
        // rows of owned atoms (each pair: half in either row), owned shares of the bonded terms, owned kinetic energy
        double v[3] = {w, h[1], 0.0};
        int rc = comm_allreduce3(c, v);
        if (rc != MC_OK) return rc;
        w = v[0]; h[1] = v[1];
    }
    if (constrained) {
        // constraint forces of the last step (SETTLE / SHAKE displacement x mass / dt^2 on the old positions)
        double hc = 0.0;
        MC_CUDA(c, cudaMemcpy(&hc, c->cons_virial.p, sizeof(hc), cudaMemcpyDeviceToHost));
        w += hc;
    }
    const double vol = (double)c->ext[0] * (double)c->ext[1] * (double)c->ext[2];
    const double ke = h[1] / (double)MC_ACCEL_CONV;
    if (virial) *virial = w;
    if (pressure_bar) *pressure_bar = (2.0 * ke + w) / (3.0 * vol) * MC_BAR_PER_KCAL_MOL_A3;
    return MC_OK;
}

extern "C" int mc_get_pressure(mc_ctx *c, double *pressure_bar, double *virial) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH_OBS(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, c->periodic, "mc_get_pressure: needs a periodic box");
    MC_REQUIRE(c, !(c->n_waters > 0 || c->n_hclusters > 0) || c->cons_virial_valid,
               "mc_get_pressure: the virial of the constraint forces is that of the last step; take a step first");
    int rc = ensure_ready(c, "mc_get_pressure");
    if (rc != MC_OK) return rc;
    if (!c->forces_valid || !c->forces_have_energy) {
        if ((rc = engine_launch_forces(c, true)) != MC_OK) return rc;
    }
    return compute_pressure(c, pressure_bar, virial);
}

// Barostat (MdConfig.barostat_cfg = BarostatCfg{pressure_target [bar], tau [ps]}, reference ui/panels/md.rs:517-556,
// properties/crystal.rs:312): every `every` steps the instantaneous pressure P of the positions just reached is
// measured and box + coordinates are scaled by mu about the origin.
//   Berendsen (weak coupling):       mu^3 = 1 - beta (every dt / tau) (P0 - P)
//   stochastic cell rescaling (Bernetti & Bussi, J. Chem. Phys. 153, 114107 (2020); the canonical companion of CSVR):
//     d(ln V) = -(beta / tau) (P0 - P) Dt + sqrt(2 kT beta Dt / (V tau)) xi,   mu = exp(d(ln V) / 3),  v <- v / mu
// beta = isothermal compressibility (1/bar).  The list is rebuilt and the forces re-evaluated on the scaled system.
static int apply_barostat(mc_ctx *c, float dt) {
    double p = 0.0;
    int rc = compute_pressure(c, &p, nullptr);
    if (rc != MC_OK) return rc;
    const double Dt = (double)dt * c->baro_every, vol = (double)c->ext[0] * c->ext[1] * c->ext[2];
    double dlnv = -(double)c->baro_beta * (Dt / (double)c->baro_tau) * ((double)c->baro_p0 - p);
    double nu = 1.0;
    if (c->baro_kind == MC_BAROSTAT_CRESCALE) {
        CsvrRng g{c->baro_seed ^ 0xB4A05747ull, c->baro_draws++, 0};
        double xi, unused;
        mc_csvr_normal_pair(g, &xi, &unused);
        const double kT = MC_KB * (double)c->lgv_temperature;
        dlnv += std::sqrt(2.0 * kT * MC_BAR_PER_KCAL_MOL_A3 * (double)c->baro_beta * Dt / (vol * (double)c->baro_tau)) * xi;
    }
    dlnv = std::max(-0.03, std::min(0.03, dlnv));  // a runaway pressure must not fold the box in one go
    const double mu = c->baro_kind == MC_BAROSTAT_CRESCALE ? std::exp(dlnv / 3.0) : std::cbrt(1.0 + dlnv);
    if (c->baro_kind == MC_BAROSTAT_CRESCALE) nu = 1.0 / mu;
    c->baro_last_p = p;
    c->baro_last_mu = mu;
    launch_scale_coords((int)c->n, c->xyzq[c->cur].p, c->xref.p, c->vel[c->cur].p, (float)mu, (float)nu, c->st, &c->launches);
    for (int a = 0; a < 3; ++a) {
        c->lo[a] = (float)((double)c->lo[a] * mu);
        c->ext[a] = (float)((double)c->ext[a] * mu);
    }
    c->grid_dirty = true;
    c->list_valid = false;
    c->forces_valid = false;
    if ((rc = engine_build_list(c)) != MC_OK) return rc;
    return engine_launch_forces(c, false);
}

extern "C" int mc_set_barostat(mc_ctx *c, int kind, float pressure_bar, float tau_ps, float compressibility_per_bar, int every_n_steps,
                               uint64_t seed) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    MC_REQUIRE(c, kind == MC_BAROSTAT_NONE || kind == MC_BAROSTAT_BERENDSEN || kind == MC_BAROSTAT_CRESCALE, "mc_set_barostat: unknown kind");
    if (kind != MC_BAROSTAT_NONE) {
        MC_REQUIRE(c, !c->comm_active, "mc_set_barostat: not available on a decomposed handle yet");
        MC_REQUIRE(c, tau_ps > 0.f && compressibility_per_bar > 0.f && every_n_steps >= 1, "mc_set_barostat: tau, compressibility and interval must be positive");
        MC_REQUIRE(c, kind != MC_BAROSTAT_CRESCALE || c->langevin || c->csvr,
                   "mc_set_barostat: stochastic cell rescaling takes its temperature from the thermostat; call mc_set_thermostat first");
    }
    c->baro_kind = kind;
    c->baro_p0 = pressure_bar; c->baro_tau = tau_ps; c->baro_beta = compressibility_per_bar; c->baro_every = every_n_steps;
    c->baro_seed = seed; c->baro_draws = 0;
    return MC_OK;
}

extern "C" int mc_get_box(mc_ctx *c, float lo[3], float hi[3]) {
    if (!c || !lo || !hi) return MC_E_INVALID;
    for (int a = 0; a < 3; ++a) { lo[a] = c->lo[a]; hi[a] = c->lo[a] + c->ext[a]; }
    return MC_OK;
}

extern "C" int mc_set_molecule_ids(mc_ctx *c, const uint16_t *mol_id) {
    if (!c) return MC_E_INVALID;
    cudaSetDevice(c->device);
    if (!mol_id) { c->have_mols = false; return MC_OK; }
    MC_CUDA(c, c->mol_of_orig.ensure((size_t)std::max<int64_t>(c->n_global, 1)));
    if (c->n_global) MC_CUDA(c, cudaMemcpy(c->mol_of_orig.p, mol_id, sizeof(uint16_t) * (size_t)c->n_global, cudaMemcpyHostToDevice));
    c->have_mols = true;
    return MC_OK;
}

extern "C" int mc_get_energy_between_mols(mc_ctx *c, double *out) {
    if (!c || !out) return MC_E_INVALID;
    MC_FLUSH_OBS(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, c->have_mols, "mc_get_energy_between_mols: call mc_set_molecule_ids first");
    int rc = ensure_ready(c, "mc_get_energy_between_mols");
    if (rc != MC_OK) return rc;
    MC_CUDA(c, c->red_out.ensure(4));
    if ((rc = engine_ensure_list32(c)) != MC_OK) return rc;
    launch_between_mols((int)c->n_rows_sorted(), (int)c->row0, c->xyzq[c->cur].p, c->type[c->cur].p, c->orig[c->cur].p, c->mol_of_orig.p,
                        c->nbr_start.p, c->nbr_count.p, c->nbr_list.p, c->ljtab.p, make_params(c), c->lj_disabled ? 0 : 1,
                        c->coul_disabled ? MC_COULOMB_NONE : c->coul_mode, c->red_out.p + 3, c->st, &c->launches);
    MC_CUDA(c, cudaMemcpyAsync(out, c->red_out.p + 3, sizeof(double), cudaMemcpyDeviceToHost, c->st));
    MC_CUDA(c, cudaStreamSynchronize(c->st));
    if (c->comm_active) {  // every rank summed the rows of its owned atoms (each pair: half in either row)
        double v[3] = {*out, 0.0, 0.0};
        if ((rc = comm_allreduce3(c, v)) != MC_OK) return rc;
        *out = v[0];
    }
    return MC_OK;
}

extern "C" int mc_get_stats(mc_ctx *c, mc_stats *out) {
    if (!c || !out) return MC_E_INVALID;
    cudaSetDevice(c->device);
    memset(out, 0, sizeof(*out));
    if (c->pairs_dirty && c->list_valid && c->n_rows_sorted() > 0) {
        // true (unpadded) entry count = sum of the row lengths
        MC_CUDA(c, c->scratch.ensure(scan_scratch_elems((size_t)c->n + 1) + 64));
        MC_CUDA(c, c->cnt_orig.ensure((size_t)c->n + 1));
        // ghost slots carry zero-length rows (cleared at build time), so the sum runs over every local slot
        exclusive_scan_u32(c->nbr_count.p, c->cnt_orig.p, (size_t)c->n, 0, c->scratch.p, c->st, &c->launches);
        uint32_t tot = 0;
        MC_CUDA(c, cudaMemcpyAsync(&tot, c->cnt_orig.p + c->n, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->st));
        MC_CUDA(c, cudaStreamSynchronize(c->st));
        c->n_pairs_listed = tot;
        c->pairs_dirty = false;
    }
    out->n_atoms = c->n_rows_sorted();
    out->n_ghosts = c->n - c->n_rows_sorted();
    out->n_pairs_listed = c->n_pairs_listed;
    out->n_rebuilds = c->n_rebuilds;
    out->n_steps = c->n_steps;
    out->n_kernel_launches = c->launches;
    for (int a = 0; a < 3; ++a) out->n_cells[a] = c->periodic ? c->h_grid.nc[a] : 0;
    out->pair_ms_sum = c->pair_acc.ms; out->pair_launches_timed = c->pair_acc.count;
    out->build_ms_sum = c->build_acc.ms; out->builds_timed = c->build_acc.count;
    out->integrate_ms_sum = c->integ_acc.ms; out->integrate_launches_timed = c->integ_acc.count;
    out->halo_ms_sum = c->halo_acc.ms; out->halos_timed = c->halo_acc.count;
    out->n_list_violations = c->n_list_violations;
    out->list_bytes = c->n_padded_entries * (int64_t)(c->list_compact ? sizeof(uint16_t) : sizeof(uint32_t));
    out->ext_upload_bytes = c->ext_upload_bytes;
    return MC_OK;
}

extern "C" int mc_reset_timers(mc_ctx *c) {
    if (!c) return MC_E_INVALID;
    c->pair_acc = c->build_acc = c->integ_acc = c->halo_acc = TimeAcc();
    return MC_OK;
}

extern "C" int mc_get_neighbors(mc_ctx *c, int64_t *start, int32_t *idx, int64_t cap, int64_t *total) {
    if (!c || !start || !total) return MC_E_INVALID;
    MC_FLUSH_OBS(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, !c->comm_active, "mc_get_neighbors: single-GPU handles only");
    MC_REQUIRE(c, c->list_valid, "mc_get_neighbors: no current list; call mc_build_neighbors");
    const int n = (int)c->n;
    if (n == 0) { start[0] = 0; *total = 0; return MC_OK; }
    MC_CUDA(c, c->cnt_orig.ensure((size_t)n + 1));
    MC_CUDA(c, c->start_orig.ensure((size_t)n + 1));
    MC_CUDA(c, c->export_rows.ensure((size_t)c->n_padded_entries + 1));
    MC_CUDA(c, c->scratch.ensure(scan_scratch_elems((size_t)n + 1) + 64));
    { int rc32 = engine_ensure_list32(c); if (rc32 != MC_OK) return rc32; }
    launch_export_rows(n, c->orig[c->cur].p, c->nbr_count.p, c->nbr_start.p, c->nbr_list.p, c->cnt_orig.p,
                       c->start_orig.p, c->export_rows.p, c->scratch.p, c->st, &c->launches);
    std::vector<uint32_t> hs((size_t)n + 1);
    MC_CUDA(c, cudaMemcpyAsync(hs.data(), c->start_orig.p, hs.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->st));
    MC_CUDA(c, cudaStreamSynchronize(c->st));
    for (int k = 0; k <= n; ++k) start[k] = hs[(size_t)k];
    *total = hs[(size_t)n];
    if (!idx) return MC_OK;
    if (cap < *total) return fail(c, MC_E_CAPACITY, "mc_get_neighbors: idx capacity too small");
    MC_CUDA(c, cudaMemcpy(idx, c->export_rows.p, (size_t)*total * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return MC_OK;
}

// ---- stand-alone kernel timing ----------------------------------------------------------------------------

extern "C" int mc_time_kernels(mc_ctx *c, int reps, int flush_l2) {
    if (!c) return MC_E_INVALID;
    MC_FLUSH(c);
    cudaSetDevice(c->device);
    MC_REQUIRE(c, reps > 0, "mc_time_kernels: reps > 0 required");
    c->prof_now = true;
    int rc = ensure_ready(c, "mc_time_kernels");
    if (rc != MC_OK) return rc;
    const bool prof = c->profiling;
    c->profiling = true;
    const size_t flush_n = flush_l2 ? (std::max<size_t>(c->l2_bytes, (size_t)128 << 20) * 2) / sizeof(float4) : 0;
    if (flush_n) MC_CUDA(c, c->flush.ensure(flush_n));
    const TimeAcc keep_pair = c->pair_acc;
    c->pair_acc = TimeAcc();
    for (int r = 0; r < reps + 3; ++r) {
        if (r == 3) { MC_CUDA(c, cudaStreamSynchronize(c->st)); c->collect_timings(); c->pair_acc = TimeAcc(); }
        if (flush_n) launch_l2_flush(c->flush.p, flush_n, c->st, &c->launches);
        if ((rc = engine_launch_forces(c, false)) != MC_OK) return rc;
    }
    MC_CUDA(c, cudaStreamSynchronize(c->st));
    c->collect_timings();
    c->last_pair_ms = c->pair_acc.count ? c->pair_acc.ms / c->pair_acc.count : 0.0;
    c->pair_acc = keep_pair;
    c->profiling = prof;
    return MC_OK;
}

extern "C" double mc_last_pair_kernel_ms(mc_ctx *c) { return c ? c->last_pair_ms : 0.0; }

// ---- docking scan -------------------------------------------------------------------------------------------

static int dock_score_impl(mc_ctx *c, int64_t n_rec, const mc_float4 *rec_xyzq, const uint16_t *rec_type,
                           const uint8_t *rec_hydrophobic, int64_t n_lig, const mc_float4 *lig_xyzq,
                           const uint16_t *lig_type, const uint8_t *lig_hydrophobic, const float lig_anchor[3],
                           int n_rec_types, int n_lig_types, const float *ljtab, int n_flex, const int32_t *flex_axis,
                           const uint8_t *flex_mask, int64_t n_poses, const float *poses, float *out);

extern "C" int mc_dock_score(mc_ctx *c, int64_t n_rec, const mc_float4 *rec_xyzq, const uint16_t *rec_type,
                             const uint8_t *rec_hydrophobic, int64_t n_lig, const mc_float4 *lig_xyzq,
                             const uint16_t *lig_type, const uint8_t *lig_hydrophobic, const float lig_anchor[3],
                             int n_rec_types, int n_lig_types, const float *ljtab, int64_t n_poses, const float *poses,
                             float *out) {
    return dock_score_impl(c, n_rec, rec_xyzq, rec_type, rec_hydrophobic, n_lig, lig_xyzq, lig_type, lig_hydrophobic, lig_anchor,
                           n_rec_types, n_lig_types, ljtab, 0, nullptr, nullptr, n_poses, poses, out);
}

extern "C" int mc_dock_score_flex(mc_ctx *c, int64_t n_rec, const mc_float4 *rec_xyzq, const uint16_t *rec_type,
                                  const uint8_t *rec_hydrophobic, int64_t n_lig, const mc_float4 *lig_xyzq,
                                  const uint16_t *lig_type, const uint8_t *lig_hydrophobic, const float lig_anchor[3],
                                  int n_rec_types, int n_lig_types, const float *ljtab, int n_flex, const int32_t *flex_axis,
                                  const uint8_t *flex_mask, int64_t n_poses, const float *poses, float *out) {
    if (!c) return MC_E_INVALID;
    MC_REQUIRE(c, n_flex >= 0 && n_flex <= MC_DOCK_MAX_FLEX && (n_flex == 0 || (flex_axis && flex_mask)),
               "mc_dock_score_flex: 0 <= n_flex <= MC_DOCK_MAX_FLEX with axis and mask arrays");
    for (int f = 0; f < n_flex; ++f)
        MC_REQUIRE(c, flex_axis[2 * f] >= 0 && flex_axis[2 * f] < n_lig && flex_axis[2 * f + 1] >= 0 && flex_axis[2 * f + 1] < n_lig &&
                          flex_axis[2 * f] != flex_axis[2 * f + 1],
                   "mc_dock_score_flex: bond atom out of range");
    return dock_score_impl(c, n_rec, rec_xyzq, rec_type, rec_hydrophobic, n_lig, lig_xyzq, lig_type, lig_hydrophobic, lig_anchor,
                           n_rec_types, n_lig_types, ljtab, n_flex, flex_axis, flex_mask, n_poses, poses, out);
}

static int dock_score_impl(mc_ctx *c, int64_t n_rec, const mc_float4 *rec_xyzq, const uint16_t *rec_type,
                           const uint8_t *rec_hydrophobic, int64_t n_lig, const mc_float4 *lig_xyzq,
                           const uint16_t *lig_type, const uint8_t *lig_hydrophobic, const float lig_anchor[3],
                           int n_rec_types, int n_lig_types, const float *ljtab, int n_flex, const int32_t *flex_axis,
                           const uint8_t *flex_mask, int64_t n_poses, const float *poses, float *out) {
    if (!c) return MC_E_INVALID;
    cudaSetDevice(c->device);
    const int stride = 7 + n_flex;
    MC_REQUIRE(c, n_rec > 0 && n_lig > 0 && n_poses >= 0 && rec_xyzq && lig_xyzq && rec_type && lig_type && ljtab &&
                      lig_anchor && (n_poses == 0 || (poses && out)),
               "mc_dock_score: NULL or empty argument");
    MC_REQUIRE(c, n_rec_types > 0 && n_lig_types > 0, "mc_dock_score: type counts must be positive");
    MC_REQUIRE(c, dock_smem_bytes((int)n_lig, n_rec_types, n_lig_types) <= 200 * 1024,
               "mc_dock_score: ligand + LJ table exceed the 200 KB shared-memory tile");
    if (n_poses == 0) return MC_OK;
    cudaStream_t st = c->st;
    std::vector<uint32_t> rm((size_t)n_rec), lm((size_t)n_lig);
    for (int64_t r = 0; r < n_rec; ++r) {
        MC_REQUIRE(c, rec_type[r] < n_rec_types, "mc_dock_score: receptor type out of range");
        rm[(size_t)r] = rec_type[r] | ((rec_hydrophobic && rec_hydrophobic[r]) ? 0x10000u : 0u);
    }
    for (int64_t a = 0; a < n_lig; ++a) {
        MC_REQUIRE(c, lig_type[a] < n_lig_types, "mc_dock_score: ligand type out of range");
        lm[(size_t)a] = lig_type[a] | ((lig_hydrophobic && lig_hydrophobic[a]) ? 0x10000u : 0u);
    }
    std::vector<float2> tab((size_t)n_rec_types * n_lig_types);
    for (size_t k = 0; k < tab.size(); ++k) tab[k] = make_float2(ljtab[2 * k] * ljtab[2 * k], 4.f * ljtab[2 * k + 1]);
    MC_CUDA(c, c->d_rec.ensure((size_t)n_rec)); MC_CUDA(c, c->d_rec_meta.ensure((size_t)n_rec));
    MC_CUDA(c, c->d_lig.ensure((size_t)n_lig)); MC_CUDA(c, c->d_lig_meta.ensure((size_t)n_lig));
    MC_CUDA(c, c->d_dock_tab.ensure(tab.size()));
    MC_CUDA(c, c->d_poses.ensure((size_t)stride * n_poses)); MC_CUDA(c, c->d_scores.ensure((size_t)5 * n_poses));
    if (n_flex > 0) {
        MC_CUDA(c, c->d_flex_axis.ensure((size_t)n_flex)); MC_CUDA(c, c->d_keep.ensure((size_t)n_flex * (size_t)n_lig));
        MC_CUDA(c, cudaMemcpyAsync(c->d_flex_axis.p, flex_axis, sizeof(int2) * (size_t)n_flex, cudaMemcpyHostToDevice, st));
        MC_CUDA(c, cudaMemcpyAsync(c->d_keep.p, flex_mask, (size_t)n_flex * (size_t)n_lig, cudaMemcpyHostToDevice, st));
    }
    MC_CUDA(c, cudaMemcpyAsync(c->d_rec.p, rec_xyzq, n_rec * sizeof(float4), cudaMemcpyHostToDevice, st));
    MC_CUDA(c, cudaMemcpyAsync(c->d_rec_meta.p, rm.data(), n_rec * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    MC_CUDA(c, cudaMemcpyAsync(c->d_lig.p, lig_xyzq, n_lig * sizeof(float4), cudaMemcpyHostToDevice, st));
    MC_CUDA(c, cudaMemcpyAsync(c->d_lig_meta.p, lm.data(), n_lig * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    MC_CUDA(c, cudaMemcpyAsync(c->d_dock_tab.p, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice, st));
    MC_CUDA(c, cudaMemcpyAsync(c->d_poses.p, poses, sizeof(float) * stride * n_poses, cudaMemcpyHostToDevice, st));
    {
        TimedRegion tr(c, c->dock_acc);
        launch_dock_score((int)n_rec, c->d_rec.p, c->d_rec_meta.p, (int)n_lig, c->d_lig.p, c->d_lig_meta.p,
                          make_float3(lig_anchor[0], lig_anchor[1], lig_anchor[2]), n_rec_types, n_lig_types,
                          c->d_dock_tab.p, (int)n_poses, c->d_poses.p, stride, n_flex, n_flex > 0 ? c->d_flex_axis.p : nullptr,
                          n_flex > 0 ? c->d_keep.p : nullptr, c->d_scores.p, st, &c->launches);
        tr.stop();
    }
    MC_CUDA(c, cudaGetLastError());
    MC_CUDA(c, cudaMemcpyAsync(out, c->d_scores.p, sizeof(float) * 5 * n_poses, cudaMemcpyDeviceToHost, st));
    MC_CUDA(c, cudaStreamSynchronize(st));
    c->collect_timings();
    c->last_dock_ms = c->dock_acc.count ? c->dock_acc.ms / c->dock_acc.count : 0.0;
    c->dock_acc = TimeAcc();
    return MC_OK;
}

// The clash pre-filter on the device (dock_filter.cu); same arguments and result as the host-only mc_dock_filter_poses.
extern "C" int mc_dock_filter_poses_gpu(mc_ctx *c, int64_t n_rec, const mc_float4 *rec_xyzq, const uint8_t *rec_is_carbon, int64_t n_lig,
                                        const mc_float4 *lig_xyzq, const uint8_t *lig_is_carbon, const float lig_anchor[3], float vdw_radius,
                                        int64_t n_poses, const float *poses, uint8_t *keep, int64_t *n_kept) {
    if (!c) return MC_E_INVALID;
    cudaSetDevice(c->device);
    MC_REQUIRE(c, rec_xyzq && rec_is_carbon && lig_xyzq && lig_is_carbon && lig_anchor && n_rec >= 0 && n_lig >= 0 && n_poses >= 0 &&
                      (n_poses == 0 || (poses && keep)),
               "mc_dock_filter_poses_gpu: NULL or negative argument");
    std::vector<float4> rs, ls;
    for (int64_t i = 0; i < n_rec; ++i)
        if (rec_is_carbon[i] && i % 6 == 0) rs.push_back(make_float4(rec_xyzq[i].x, rec_xyzq[i].y, rec_xyzq[i].z, 0.f));
    for (int64_t i = 0; i < n_lig; ++i)
        if (lig_is_carbon[i] && i % 4 == 0) ls.push_back(make_float4(lig_xyzq[i].x, lig_xyzq[i].y, lig_xyzq[i].z, 0.f));
    MC_REQUIRE(c, (int)ls.size() <= dock_filter_max_lig(), "mc_dock_filter_poses_gpu: more than " + std::to_string(dock_filter_max_lig()) +
                                                              " sampled ligand carbons");
    if (n_poses == 0) { if (n_kept) *n_kept = 0; return MC_OK; }
    cudaStream_t st = c->st;
    MC_CUDA(c, c->d_rec_s.ensure(std::max<size_t>(rs.size(), 1))); MC_CUDA(c, c->d_lig_s.ensure(std::max<size_t>(ls.size(), 1)));
    MC_CUDA(c, c->d_poses.ensure((size_t)7 * n_poses)); MC_CUDA(c, c->d_keep.ensure((size_t)n_poses));
    if (!rs.empty()) MC_CUDA(c, cudaMemcpyAsync(c->d_rec_s.p, rs.data(), rs.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
    if (!ls.empty()) MC_CUDA(c, cudaMemcpyAsync(c->d_lig_s.p, ls.data(), ls.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
    MC_CUDA(c, cudaMemcpyAsync(c->d_poses.p, poses, sizeof(float) * 7 * n_poses, cudaMemcpyHostToDevice, st));
    launch_dock_filter((int)rs.size(), c->d_rec_s.p, (int)ls.size(), c->d_lig_s.p, make_float3(lig_anchor[0], lig_anchor[1], lig_anchor[2]),
                       vdw_radius * 1.1f, (int)n_poses, c->d_poses.p, c->d_keep.p, st, &c->launches);
    MC_CUDA(c, cudaGetLastError());
    MC_CUDA(c, cudaMemcpyAsync(keep, c->d_keep.p, (size_t)n_poses, cudaMemcpyDeviceToHost, st));
    MC_CUDA(c, cudaStreamSynchronize(st));
    if (n_kept) {
        int64_t k = 0;
        for (int64_t p = 0; p < n_poses; ++p) k += keep[p] ? 1 : 0;
        *n_kept = k;
    }
    return MC_OK;
}

extern "C" double mc_last_dock_kernel_ms(mc_ctx *c) { return c ? c->last_dock_ms : 0.0; }
