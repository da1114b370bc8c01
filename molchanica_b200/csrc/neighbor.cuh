// neighbor.cuh -- host-side launchers of neighbor.cu
#pragma once
#include "common.cuh"

struct ReorderArrays {
    const float4 *xyzq_in; float4 *xyzq_out; float4 *xref;
    const float4 *vel_in; float4 *vel_out;
    const uint16_t *type_in; uint16_t *type_out;
    const uint8_t *flags_in; uint8_t *flags_out;
    const int *orig_in; int *orig_out; int *slot_of_orig;
    uint32_t *cell_start;
    int mark_interior;  // set MC_FLAG_INTERIOR on atoms whose stencil never wraps
};

void launch_bbox(const float4 *xyzq, int n, float *bb, float cw_min, int max_cells, GridParams *g, cudaStream_t st,
                 int64_t *launches);
void launch_wrap_key(float4 *xyzq, int n, const GridParams *g, uint32_t *keys, uint32_t *vals, cudaStream_t st,
                     int64_t *launches);
void launch_reorder(int n, const uint32_t *skeys, const uint32_t *svals, const GridParams *g, const ReorderArrays &a,
                    cudaStream_t st, int64_t *launches);
void launch_sweep(bool fill, int n_rows, const float4 *xyzq, const uint32_t *cell_start, const GridParams *g, float rl2,
                  const uint32_t *cell_of_slot, const int *orig, const int32_t *excl_start, const int32_t *excl_idx, uint32_t *nbr_count,
                  const uint32_t *nbr_start, uint32_t *nbr_list, cudaStream_t st, int64_t *launches);
void launch_export_rows(int n, const int *orig, const uint32_t *nbr_count, const uint32_t *nbr_start,
                        const uint32_t *nbr_list, uint32_t *cnt_orig, uint32_t *start_orig, uint32_t *rows,
                        uint32_t *scan_scratch, cudaStream_t st, int64_t *launches);

// tile_build.cu -- single-pass TMA-staged build
cudaError_t tile_sweep_prepare();
uint32_t tile_sweep_max_atoms();
uint32_t tile_sweep_max_atoms_dense();  // rows_build_kernel, one CTA per SM (dense systems)
void launch_tile_build(int n_rows, int grid_cells, int split, int n_sms, const float4 *xyzq, const uint32_t *cell_start,
                       const GridParams *g, float rl2, float rc2_inner, const int *orig, const int32_t *excl_start,
                       const int32_t *excl_idx, uint32_t *nbr_count, uint32_t *nbr_start, void *nbr_list, bool compact, bool partition,
                       uint32_t list_cap, uint32_t tile_cap, uint32_t *ctl, cudaStream_t st, int64_t *launches, int variant = 1,
                       uint32_t row_hint = 0, uint32_t *plan = nullptr, int variant_min_blocks = 3,
                       const int *slot_of_orig = nullptr /* rows_build_kernel: exclusions filtered after the sweep */);
size_t rows_plan_words(int grid_cells);  // size of `plan` (uint32_t), the per-cell staging records of variant 2
// variant = 2 (the engine's default; needs `plan`): global-slot rows without partition are built by rows_build_kernel (ballot compaction, rows staged in
// shared memory when row_hint -- the longest row of the previous build, ctl[6] -- says they fit); 1: tile_build_kernel always.
// compact = true: nbr_list is uint16_t[list_cap], entries are tile-local indices (tile_ring.cuh) -- what pair_tile.cu reads;
// compact = false: uint32_t[list_cap] global slots.  Rows are padded to 8 entries either way.
// partition = true: entries inside the force cutoff at build time come first in every row (the warp-uniform pair loop skips
// the skin shell warp-wide); false (default): plain tile order, cheaper to build.
// Compact rows -> global-slot rows with the same nbr_start / nbr_count (for the consumers that are not on the hot path)
void launch_expand_rows(int grid_cells, const uint32_t *cell_start, const GridParams *g, const uint32_t *nbr_start,
                        const uint32_t *nbr_count, const uint16_t *list16, uint32_t *list32, cudaStream_t st, int64_t *launches);
