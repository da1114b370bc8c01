// engine.cuh -- the handle behind include/molchanica_md.h (shared by engine.cu and comm.cu)
#pragma once
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "common.cuh"
#include "halo_sync.cuh"
#include "bonded.cuh"
#include "settle.cuh"
#include "pme.cuh"
#include "group_energy.cuh"

// Decomposed handles size many buffers by what a rank holds (owned + ghost atoms), which drifts by a few atoms per rebuild: with
// exact first allocations every such buffer is freed and allocated again the first time the count goes up -- possibly hundreds
// of steps into a run, and a cudaFree is a device-wide synchronisation (seen as sporadic 40-190 ms steps in the 2-GPU
// end-to-end leg, profiles/e2e_stall_r2w1.txt).  mc_comm_init sets this: first allocations get the head-room as well.
inline bool g_devbuf_roomy = false;

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    // grow-only; contents are NOT preserved across a growth
    cudaError_t ensure(size_t want) {
        if (want <= n && p) return cudaSuccess;
        // a buffer that has to grow a second time gets head-room: sizes that follow the atoms a rank owns drift up and down by a
        // few atoms per rebuild, and every cudaFree is a device-wide synchronisation (with a collective in flight: milliseconds)
        static const bool trace = getenv("MC_TRACE_ALLOC") != nullptr;  // stderr line per (re)allocation: finds stalls of a steady loop
        if (trace) fprintf(stderr, "[mc alloc] %zu -> %zu elements of %zu bytes%s\n", n, want, sizeof(T), p ? " (cudaFree + cudaMalloc)" : "");
        if (p || g_devbuf_roomy) want += want / 8 + 256;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        if (want == 0) want = 1;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) n = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

struct TimeAcc {
    double ms = 0.0;
    int64_t count = 0;
};

struct EventPair {
    cudaEvent_t a, b;
};

struct CommState;  // comm.cu
struct HaloSplit { int n_first, last_begin; HaloWait wait; };

// what an mc_step call still has to look at once its kernels have finished (engine.cu: step_epilogue)
struct StepEpilogue {
    bool pending = false;  // a pipelined call returned without synchronising: the next call / the closer of the open step collects
    bool pipelined = false, skip_prev = false, check_flag = false, flags_arrive = false, trace_dev = false;
    int n_steps = 0, n_ranks_f = 1;
};

struct mc_ctx {
    int device = 0;
    int n_sms = 148;
    size_t l2_bytes = 0;
    cudaStream_t st = nullptr;
    std::string err;
    void *h_pinned = nullptr;
    int *h_flags_all = nullptr;   // pinned, 2 ints per rank: the flag words that ride with the external-force all-gather of a pipelined decomposed call
    StepEpilogue epi;
    cudaEvent_t ev_drift = nullptr;  // recorded right after the last drift of a pipelined call
    bool drift_event_valid = false;
    // Two ways to end a pipelined call, measured on B200s (profiles/e2e_r2_options.txt) and left OFF: both lose to the plain
    // synchronised ending because the per-step loop is bound by its two PCIe transfers, which the plain ending already overlaps
    // perfectly (the snapshot copy of step k under the upload of step k+1)
    bool early_tail = false;      // option "early_tail": a pipelined call launches the open step's force evaluation before it returns
    bool lazy_sync = false;       // option "lazy_sync": pipelined calls (external forces, defer_tail) return without waiting for their kernels
    bool flags_ride = false;      // the previous call was such a call: its flags arrive with this call's all-gather
    // MC_TRACE_STEP=1: host time spent in the phases of mc_step (seconds, summed; printed by mc_destroy) -- a debugging aid
    bool trace_step = false;
    double trace_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t trace_calls = 0;
    cudaEvent_t trace_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double trace_dev[4] = {0, 0, 0, 0};  // device time of {upload, tail (forces + half kick), all-gather, kick + drift} of the traced calls (ms, summed)

    // system
    int64_t n = 0;         // atoms held locally (owned + ghosts)
    int64_t n_rows = 0;    // atoms owned locally (rows of the list); == n on a single GPU
    int64_t row0 = 0;      // first owned slot of the cell-ordered arrays (ghost layer in front when decomposed)
    int64_t n_global = 0;  // atoms of the whole system (original ids run over this range)
    bool periodic = false;
    float lo[3] = {0, 0, 0}, ext[3] = {1, 1, 1};
    float rc_lj = 0, rc_q = 0, skin = 0, alpha = 0.35f;
    int coul_mode = MC_COULOMB_NONE;
    bool lj_disabled = false, coul_disabled = false;
    int n_types = 0;
    float sig2_0 = 0, eps24_0 = 0;
    float scale14_lj = 0.5f, scale14_q = 1.0f / 1.2f;
    bool have_excl = false, have_p14 = false;

    // state flags
    bool forces_have_energy = false;
    bool grid_dirty = true, list_valid = false, forces_valid = false, identity_order = true, pairs_dirty = true;
    int cur = 0;
    const uint32_t *cell_of_slot = nullptr;  // sorted cell keys of the current build (slot -> local cell)
    int key_bits = 1;
    size_t ncell_cap = 0;
    float cw_min = 0;
    GridParams h_grid{};
    int pair_lanes = 8;
    bool pair_uniform = false;  // warp-uniform pair loop (skips warp-wide skin-shell iterations)
    int rebuild_every = 0;
    int steps_since_build = 0;
    bool profiling = false;
    int prof_every = 1;     // option "profile_every": inside mc_step only every k-th step's kernels are bracketed
    bool prof_now = true;   // (event pairs between back-to-back kernels cost ~10 us per step on a 0.2 ms step)
    bool subcell_sort = false;  // Morton sub-cell code in the low sort-key bits (option "subcell_sort")
    bool use_tile = true;      // TMA-staged tile sweep for the list build (neighbor_tile.cu)
    uint32_t tile_cap = 1024;  // tile capacity in atoms, grows on demand
    int pair_tile_stages = 0;  // option "pair_tile_stages": tiles in flight per CTA of pair_tile.cu (0 = from the shared-memory budget)
    int rows_interleave = 0;   // option "rows_interleave": keep a quad-interleaved copy of the rows for the 8-lane force kernel
                               // (measured on C4, profiles/rows_interleave_r2s1.txt: pair kernel 0.163 vs 0.1585 ms with plain rows
                               // and +0.15 ms per rebuild -- the index loads are not what occupies the L1 data pipe; default off)
    bool ilv_valid = false;    // nbr_ilv / ilv_qbase describe the current list
    int build_variant = 2;     // option "build_variant": 2 = rows_build_kernel (default), 1 = tile_build_kernel (tile_build.cu)
    uint32_t row_len_hint = 0; // longest row of the last build (0: none yet)
    int build_split = 0;       // option "build_split": slices per cell of the list build (0 = from the cell count)
    bool rows_dense = true;    // option "rows_dense": 16 consumer warps + single direct sweep where a tile leaves room for one CTA per SM
    int rows_min_blocks = 3;   // option "rows_min_blocks"
    uint32_t row_stage_limit = 0;  // option "row_stage_limit" (testing): cap on the hint, forces the two-sweep path for longer rows
    int use_pair_tile = 0;     // option "pair_tile": compact rows (16-bit tile-local indices) + TMA-staged force kernel (pair_tile.cu):
                               // 0 (default) off, 1 on, 2 on for systems of >= 16384 atoms.  Measured on C4 (profiles/pair_tile_r2_*):
                               // 0.199 ms against 0.178 ms of the gather kernel -- staging a 27-cell tile per ~19-atom cell is
                               // latency bound (0.126 ms with the arithmetic removed), so it is an option, not the default
    bool pair_tile_fits = true; // cleared when a build showed the system too dense for pair_tile.cu (cells of > 32 atoms, rows too long)
    bool list_compact = false; // the current list is in compact form (nbr_list16)
    bool list32_valid = false; // nbr_list holds the current list as global slots (expanded on demand from the compact form)
    uint32_t tile_max_m = 0;   // largest tile (atoms) of the current build = stage capacity of the force kernel
    uint32_t rows_max_entries = 0;  // largest row block (16-bit entries) one 32-atom pass of the build allocated

    // counters
    int64_t n_list_violations = 0;
    int64_t ext_upload_bytes = 0;  // H2D bytes of the last mc_step's external forces
    int64_t launches = 0, n_rebuilds = 0, n_steps = 0, n_pairs_listed = 0, n_padded_entries = 0;
    TimeAcc pair_acc, build_acc, integ_acc, halo_acc, dock_acc;
    double last_pair_ms = 0, last_dock_ms = 0, last_step_ms = 0;
    cudaEvent_t ev_step_a = nullptr, ev_step_b = nullptr, ev_flag[2] = {nullptr, nullptr};
    int flag_tag = 0;           // tag of the last published rebuild-flag word (kick_drift -> pinned host memory)
    bool sync_rebuild = false;  // poll the displacement flag synchronously every step (debug / comparison)

    // device arrays
    DevBuf<float4> xyzq[2], vel[2], force, xref, stage, flush;
    DevBuf<uint16_t> type[2];
    DevBuf<uint8_t> flags[2];
    DevBuf<int> orig[2], slot_of_orig, rebuild_flag;
    DevBuf<uint32_t> keys[2], vals[2], scratch, cell_start, nbr_count, nbr_start, nbr_list;
    DevBuf<uint32_t> nbr_ilv, ilv_qbase;  // quad-interleaved copy of the rows for the 8-lane force kernel (pair_force.cu)
    DevBuf<uint32_t> cnt_orig, start_orig, export_rows, tile_need, cell_plan, cell_rowtab, rows_plan;
    DevBuf<uint16_t> nbr_list16;
    DevBuf<int32_t> excl_start, excl_idx, p14_start, p14_idx;
    DevBuf<float2> ljtab, d_dock_tab;
    DevBuf<float> bbox, ext_force, d_poses, d_scores;
    DevBuf<GridParams> grid;
    DevBuf<double> red_partial, red_out;
    DevBuf<float4> d_rec, d_lig, d_rec_s, d_lig_s;
    DevBuf<uint8_t> d_keep;
    DevBuf<int2> d_flex_axis;
    DevBuf<uint32_t> d_rec_meta, d_lig_meta;

    // bonded terms (bonded.cu), caller's atom ids
    int n_bonds = 0, n_angles = 0, n_dihedrals = 0;
    DevBuf<int2> bonds;
    DevBuf<float2> bond_kr0, angle_kt0;
    DevBuf<int4> angles, dihedrals;
    DevBuf<float4> dihedral_prm;
    // barostat (mc_set_barostat): every baro_every steps the box and all coordinates are scaled towards the target pressure
    int baro_kind = 0, baro_every = 10;
    float baro_p0 = 1.f, baro_tau = 5.f, baro_beta = 4.5e-5f;
    uint64_t baro_seed = 0, baro_draws = 0;
    double baro_last_p = 0.0, baro_last_mu = 1.0;
    int com_every = 0;            // option "zero_com_drift": remove the centre-of-mass velocity every k steps (0 = never)
    DevBuf<double> com_partial;
    DevBuf<double> cons_virial;   // virial of the constraint forces of the last step (settle.cu)
    bool cons_virial_valid = false;
    DevBuf<double> bonded_e;   // {E_bond, E_angle, E_dihedral} of the last evaluation that asked for energies
    DevBuf<float4> min_x, min_v;   // energy minimiser: positions / velocities saved in original order
    bool have_mols = false;    // molecule ids for energy_potential_between_mols (group_energy.cu)
    DevBuf<uint16_t> mol_of_orig;
    bool csvr = false;         // CSVR thermostat (thermostat.cu); lgv_gamma then holds 1 / tau
    DevBuf<float> csvr_lambda;
    bool langevin = false;     // Langevin thermostat (integrate.cu langevin_ou_kernel)
    float lgv_temperature = 300.f, lgv_gamma = 1.f;
    uint64_t lgv_seed = 0, lgv_step = 0;
    PmeState pme;              // SPME reciprocal space (pme.cu); pme.planned == false: off
    int n_waters = 0;          // rigid three-site waters (settle.cu)
    DevBuf<int4> waters;
    float water_m_o = 0, water_m_h = 0, water_d_oh = 0, water_d_hh = 0;
    int n_hclusters = 0, n_hconstraints = 0;   // SHAKE clusters of bonds to hydrogen (settle.cu)
    DevBuf<int4> hclusters;
    DevBuf<float> hdist;
    // host-side record of which constraint owns an atom (original ids): 1 = a rigid water, 2 = a hydrogen cluster.  One thread
    // owns a molecule / cluster and writes its atoms without atomics, so no atom may sit in two of them
    std::vector<uint8_t> h_in_water, h_in_hcluster;
    DevBuf<int> shake_fail;
    DevBuf<int> bonded_missing;  // decomposed handles: set by bonded_kernel when a term's partner is not held by this rank
    float shake_tol = 1e-6f;
    int n_vsites = 0;          // virtual sites of four-site water (settle.cu)
    DevBuf<int4> vsites;
    float vsite_a = 0, vsite_b = 0;
    double total_mass = 0.0;   // amu, from the inverse masses handed to mc_set_atoms (density of the snapshot)

    // pipelined external forces (engine.cu, mc_step): the caller's array is uploaded on a stream of its own while the
    // force evaluation the previous call left open runs; that call's second half kick is applied once both are there
    int fused_lanes = 0;         // option "fused_lanes": lanes per row in the fused kernel's force phase (0 = chosen by the launcher)
    bool fused_brute = true;     // option "fused_brute": up to md_fused_brute_max_atoms() atoms the fused kernel keeps its own all-pairs list
    DevBuf<uint32_t> bl_list, bl_count, bl_flags;  // the private list of the fused kernel's brute-force mode
    DevBuf<float4> bl_xref;
    DevBuf<unsigned long long> fused_dbg;
    bool bl_valid = false;       // the private list matches the atoms as they are (positions moved only by the fused kernel since)
    int64_t bl_n = 0;
    float bl_rl2 = 0.f;
    bool fused_steps = true;     // option "fused_steps": small plain-NVE systems take all steps of a call in one cooperative launch (md_fused.cu)
    bool defer_tail = false;     // option "defer_tail" (opt-in; confirmed on hardware in round 2, bench.py's e2e leg turns it on)
    bool tail_pending = false;   // positions are one step ahead of forces / velocities (half kick outstanding)
    float tail_dt = 0.f;
    const float *tail_ext = nullptr;  // external forces of the outstanding half kick (device, one of the two buffers)
    bool tail_rebuild = false;        // decomposed: the open step ends in a scheduled rebuild (before its force evaluation)
    bool tail_use_split = false;      // decomposed, fused halo: the open force evaluation waits on the ready flags of ...
    HaloSplit tail_split{};           // ... this step
    DevBuf<float> ext_force2;
    int ext_k = 0;
    cudaStream_t st_up = nullptr;
    cudaEvent_t ev_up = nullptr;

    // asynchronous snapshots (mc_snapshot_begin / mc_snapshot_wait): double-buffered staging + a copy stream
    cudaStream_t st_copy = nullptr;
    cudaEvent_t ev_snap_staged[2] = {nullptr, nullptr}, ev_snap_done[2] = {nullptr, nullptr};
    DevBuf<float4> snap_stage[2], snap_stage_v[2];
    DevBuf<int> snap_ids[2];
    int snap_k = 0;
    bool snap_pending[2] = {false, false};

    // event timing
    std::vector<EventPair> ev_pool;
    struct Pending { int ev; TimeAcc *acc; };
    std::vector<Pending> pending;
    size_t ev_used = 0;

    // domain decomposition
    bool comm_active = false;
    bool halo_fused = true;  // option "halo_fused": peer-memory halo inside the step kernels (else NCCL send / recv)
    CommState *comm = nullptr;

    int64_t n_rows_sorted() const { return n_rows; }

    cudaError_t alloc_atoms(size_t m) {
        cudaError_t e;
        const size_t cap = m + 8;
        for (int b = 0; b < 2; ++b) {
            if ((e = xyzq[b].ensure(cap)) != cudaSuccess) return e;
            if ((e = vel[b].ensure(cap)) != cudaSuccess) return e;
            if ((e = type[b].ensure(cap)) != cudaSuccess) return e;
            if ((e = flags[b].ensure(cap)) != cudaSuccess) return e;
            if ((e = orig[b].ensure(cap)) != cudaSuccess) return e;
        }
        if ((e = force.ensure(cap)) != cudaSuccess) return e;
        if ((e = xref.ensure(cap)) != cudaSuccess) return e;
        if ((e = slot_of_orig.ensure(std::max<size_t>(cap, (size_t)n_global + 8))) != cudaSuccess) return e;
        if ((e = nbr_count.ensure(cap)) != cudaSuccess) return e;
        if ((e = nbr_start.ensure(cap + 1)) != cudaSuccess) return e;
        if ((e = rebuild_flag.ensure(4)) != cudaSuccess) return e;
        return cudaMemset(rebuild_flag.p, 0, 4 * sizeof(int));
    }

    void free_all() {
        for (int b = 0; b < 2; ++b) { xyzq[b].release(); vel[b].release(); type[b].release(); flags[b].release(); orig[b].release(); keys[b].release(); vals[b].release(); }
        fused_dbg.release(); bl_list.release(); bl_count.release(); bl_flags.release(); bl_xref.release();
        force.release(); xref.release(); stage.release(); flush.release(); slot_of_orig.release(); rebuild_flag.release();
        scratch.release(); cell_start.release(); nbr_count.release(); nbr_start.release(); nbr_list.release();
        nbr_ilv.release(); ilv_qbase.release();
        cnt_orig.release(); start_orig.release(); export_rows.release(); tile_need.release(); rows_plan.release(); cell_plan.release(); cell_rowtab.release(); nbr_list16.release();
        excl_start.release(); excl_idx.release(); p14_start.release(); p14_idx.release();
        ljtab.release(); d_dock_tab.release(); bbox.release(); ext_force.release(); ext_force2.release(); d_poses.release(); d_scores.release();
        grid.release(); red_partial.release(); red_out.release(); d_rec.release(); d_lig.release();
        d_rec_meta.release(); d_lig_meta.release(); d_rec_s.release(); d_lig_s.release(); d_keep.release(); d_flex_axis.release();
        for (int b = 0; b < 2; ++b) { snap_stage[b].release(); snap_stage_v[b].release(); snap_ids[b].release(); }
        bonds.release(); bond_kr0.release(); angle_kt0.release(); angles.release(); dihedrals.release(); dihedral_prm.release();
        bonded_e.release(); cons_virial.release(); com_partial.release(); waters.release(); vsites.release(); csvr_lambda.release();
        hclusters.release(); hdist.release(); shake_fail.release(); bonded_missing.release(); mol_of_orig.release(); min_x.release(); min_v.release();
    }

    // resolve the CUDA-event pairs recorded since the last call (stream must be idle)
    void collect_timings() {
        for (const Pending &p : pending) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev_pool[(size_t)p.ev].a, ev_pool[(size_t)p.ev].b) == cudaSuccess) {
                p.acc->ms += ms;
                p.acc->count += 1;
            }
        }
        pending.clear();
        ev_used = 0;
    }
};

// Brackets kernel launches with CUDA events on the handle's stream when profiling is on.
struct TimedRegion {
    mc_ctx *c;
    TimeAcc &acc;
    int ev = -1;
    TimedRegion(mc_ctx *c_, TimeAcc &a, bool sampled = false) : c(c_), acc(a) {
        if (!c->profiling || c->ev_used >= 8192 || (sampled && !c->prof_now)) return;
        if (c->ev_used >= c->ev_pool.size()) {
            EventPair p;
            if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return;
            c->ev_pool.push_back(p);
        }
        ev = (int)c->ev_used++;
        cudaEventRecord(c->ev_pool[(size_t)ev].a, c->st);
    }
    void stop() {
        if (ev < 0) return;
        cudaEventRecord(c->ev_pool[(size_t)ev].b, c->st);
        c->pending.push_back({ev, &acc});
        ev = -1;
    }
    ~TimedRegion() { stop(); }
};

// engine.cu
int engine_build_list(mc_ctx *c);
int engine_flush_tail(mc_ctx *c);  // closes the half kick mc_step left open (pipelined external forces)
int engine_build_rows(mc_ctx *c);
int engine_ensure_list32(mc_ctx *c);  // global-slot rows for the consumers off the hot path (expands the compact list once per build)
// hs != nullptr: decomposed step with the peer-memory halo -- interior rows first, then the rows of the
// first and last owned layer in one launch that waits for the neighbours' pushes
int engine_launch_forces(mc_ctx *c, bool want_energy, const HaloSplit *hs = nullptr);
int engine_upload_local(mc_ctx *c, int64_t n, const mc_float4 *xyzq, const uint16_t *type, const mc_float4 *vel,
                        const uint8_t *flags, const int *orig_ids, size_t alloc_n);

// comm.cu -- slab domain decomposition + ghost-atom halo exchange
int comm_set_atoms(mc_ctx *c, int64_t n, const mc_float4 *xyzq, const uint16_t *type, const mc_float4 *vel,
                   const uint8_t *flags);
int comm_rebuild(mc_ctx *c);          // migrate + re-select ghosts + engine_build_list
int comm_halo_positions(mc_ctx *c);   // per-step ghost position refresh (NCCL send / recv)
bool comm_peer_direct(const mc_ctx *c);
void comm_set_migrate(mc_ctx *c, bool on);
int comm_interval(const mc_ctx *c);  // steps between two builds of a decomposed run  // the fused peer-memory halo is usable on this communicator
// one decomposed step with the fused halo: advances the epoch and fills the push descriptor of this
// step's kick_drift (no push when the step ends in a rebuild) and the wait descriptor of its pair kernels
void comm_step_descriptors(mc_ctx *c, bool rebuild_step, HaloPush *push, HaloSplit *split);
int comm_agree_flag(mc_ctx *c, bool *flag);
int comm_reduce_flags_async(mc_ctx *c, const int *d_flags2, int *h_out2);
void comm_shrink_interval(mc_ctx *c);
void comm_first_interval(mc_ctx *c, float dt);  // first adaptive interval from the fastest atom and the time step (comm.cu)
int comm_allreduce3(mc_ctx *c, double v[3]);
int comm_allreduce_dev_f64(mc_ctx *c, double *d, int n);  // in place, device memory, engine stream, no host sync
int comm_allreduce_f4(mc_ctx *c, float4 *buf, int64_t n);  // in-place sum over ranks
void comm_rank_size(const mc_ctx *c, int *rank, int *n_ranks);
int comm_allgather_ext_and_flags(mc_ctx *c, float *buf, size_t chunk, const int *d_flags2, int *h_flags);  // h_flags: pinned, 2 ints per rank
int comm_allgather_f32_inplace(mc_ctx *c, float *buf, size_t chunk);  // rank r contributes buf[r * chunk .. (r + 1) * chunk)
void comm_destroy(mc_ctx *c);
