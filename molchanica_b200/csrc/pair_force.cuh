// pair_force.cuh -- host-side launchers of pair_force.cu
#pragma once
#include "common.cuh"
#include "halo_sync.cuh"

struct PairLaunch {
    int n_rows;
    int row0;  // first row slot (rows are slots row0 .. row0 + n_rows - 1)
    const float4 *xyzq;
    const uint16_t *type;
    const uint8_t *flags;  // MC_FLAG_INTERIOR decides whether a row needs the minimum image
    const uint32_t *nbr_start, *nbr_count, *nbr_list;
    const float2 *ljtab;  // T*T (sigma^2, 24 eps)
    NbParams p;
    int lj_on;
    int coul;   // MC_COULOMB_*
    bool multi; // more than one LJ type
    int lanes;  // lanes per row: 4, 8, 16 or 32
    bool energy;  // also accumulate the per-atom energy row sum into force.w
    bool uniform; // warp-uniform row loop with warp-wide skin-shell skipping
    float4 *force;
    // decomposed rank with the fused halo: launch rows run interior rows first (n_interior of them, slots
    // row0 + n_first ...), then the first owned layer (n_first rows) and the last one; blocks past the
    // interior rows wait on the neighbours' ready flags.  Defaults = plain order, no wait.
    int n_interior = 0, n_first = 0;
    HaloWait wait{};
    // quad-interleaved copy of the rows (launch_rows_interleave_*), used by the 8-lane kernel when set
    const uint32_t *ilv_qbase = nullptr, *ilv_list = nullptr;
};

int pair_force_max_types();
// Quad-interleaved rows: sizes (qbase per quad of four slots, *cursor += entries claimed) then copy; see pair_force.cu.
void launch_rows_interleave_sizes(int n_slots, const uint32_t *nbr_count, uint32_t *qbase, uint32_t *cursor, cudaStream_t st, int64_t *launches);
void launch_rows_interleave_copy(int n_slots, const uint32_t *nbr_start, const uint32_t *nbr_count, const uint32_t *nbr_list,
                                 const uint32_t *qbase, uint32_t *out, cudaStream_t st, int64_t *launches);
cudaError_t pair_force_prepare();
void launch_pair_force(const PairLaunch &L, cudaStream_t st, int64_t *launches);
void launch_pairs14(int n_rows, int row0, const float4 *xyzq, const uint16_t *type, const int *orig, const int *slot_of_orig,
                    const int32_t *p14_start, const int32_t *p14_idx, const float2 *ljtab, const NbParams &p,
                    float scale_lj, float scale_q, int lj_on, int coul_on, float4 *force, cudaStream_t st,
                    int64_t *launches);
int energy_partial_elems();
void launch_energy_reduce(int n_rows, const float4 *force, const float4 *vel, const uint8_t *flags /* may be NULL */, double *partial, double *out3,
                          cudaStream_t st, int64_t *launches);

// pair_tile.cu -- the TMA-staged variant: rows of 16-bit tile-local indices (tile_build.cu, compact = true), the cell's
// 27-cell tile in shared memory.  ctl: 4 zeroed words owned by this launcher (the kernel re-arms them itself).
struct PairTileLaunch {
    int grid_cells;   // cells of the (periodic) grid
    int n_sms;
    const float4 *xyzq;
    const uint16_t *type;
    const GridParams *grid;
    const uint32_t *plan, *rowtab;  // launch_cell_plan, once per build
    const uint16_t *list16;
    const float2 *ljtab;
    NbParams p;
    int lj_on, coul;
    bool multi, energy;
    float4 *force;
    uint32_t tile_cap;
    uint32_t rows_max_entries;  // largest row block of a cell
    int force_stages = 0;       // option pair_tile_stages (0: chosen from the shared-memory budget)
    HaloWait wait{};
};
cudaError_t pair_tile_prepare();
size_t pair_tile_smem(uint32_t tile_cap, uint32_t rows_max_entries, int n_types, bool multi, int *n_stages_out, uint32_t *rows_cap_out,
                      int force_stages = 0);
void launch_pair_tile(const PairTileLaunch &L, cudaStream_t st, int64_t *launches);
// per-build table of everything the force kernel's producer needs per cell (plan: n_cells x pair_tile_plan_words() words,
// rowtab: n_cells x 32 words); ctl[3] is set to 2 when a cell cannot be staged (> 32 atoms, rows not one block, ...)
size_t pair_tile_plan_words();
void launch_cell_plan(int n_cells, const uint32_t *cell_start, const GridParams *g, const uint32_t *nbr_start, const uint32_t *nbr_count,
                      uint32_t tile_cap, uint32_t rows_cap, uint32_t *plan, uint32_t *rowtab, uint32_t *ctl, cudaStream_t st, int64_t *launches);
