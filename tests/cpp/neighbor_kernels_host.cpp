// Test infrastructure: the cell-list kernels of neighbor.cu (wrap + cell key, reorder + cell starts + interior flag,
// the 27-cell sweep with its exact accept test, the vacuum bounding-box grid) compiled unchanged for the host through
// tests/cpp/shim_mt/cuda_runtime.h and chained the way engine.cu chains them; with the radix sort and the prefix
// scan of sort_scan.cu (kernels and host drivers) in between, exactly the sequence of engine_build_list up to the list sweep.  tests/test_neighbor_kernels_on_host.py holds the resulting Verlet
// list to the oracle's, index for index.  With use_tile the rows come from
// the production list build instead (tile_build.cu through tests/cpp/tile_build_host.cpp), driven by the adaptive loop of
// engine_build_rows.
#define MC_HOST_SHIM 1
#define MC_SHIM_SHARED_STATIC 1
#include "shim_mt/cuda_runtime.h"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "../../molchanica_b200/csrc/neighbor.cu"
#include "../../molchanica_b200/csrc/sort_scan.cu"
#include "../../molchanica_b200/csrc/neighbor.cuh"  // launch_tile_build: tests/cpp/tile_build_host.cpp

extern "C" {

// xyzq: n x 4 (original order).  Returns the number of list entries; fills (all in CELL order, capacity given by the caller):
// orig_out[n], flags_out[n], xyzq_sorted[n x 4], nbr_count[n], nbr_start[n] (rows padded to 8), nbr_list[cap].
// n_cells_out[3] reports the grid.  -1: capacity too small.
long host_neighbor_list(int n, const float4 *xyzq_in, const float *lo, const float *ext, int periodic, float r_list,
                        const int32_t *excl_start, const int32_t *excl_idx, int *orig_out, uint8_t *flags_out, float4 *xyzq_sorted,
                        uint32_t *nbr_count, uint32_t *nbr_start, uint32_t *nbr_list, long cap, int *n_cells_out,
                        int use_tile, float rc_inner, int tile_cap0, int list_cap0, int split, int n_sms, int *tile_stats) {
    std::vector<float4> x(xyzq_in, xyzq_in + n), xo(n), xref(n), vin(n, float4{0, 0, 0, 1}), vout(n);
    std::vector<uint16_t> tin(n, 0), tout(n);
    std::vector<uint8_t> fin(n, 0);
    std::vector<int> oin(n), slot_of_orig(n);
    std::iota(oin.begin(), oin.end(), 0);
    GridParams g;
    memset(&g, 0, sizeof(g));
    const double cw_min = (double)r_list * 1.001 + 1e-3;  // engine.cu setup_grid
    size_t ncell_cap;
    if (periodic) {
        long long ncell = 1;
        for (int a = 0; a < 3; ++a) {
            int m = (int)std::floor((double)ext[a] / cw_min);
            m = std::max(1, std::min(m, 1024));
            g.nc[a] = m; g.lo[a] = lo[a]; g.ext[a] = ext[a];
            g.inv_ext[a] = 1.0f / ext[a];
            g.inv_cw[a] = (float)((double)m / (double)ext[a]);
            ncell *= m;
        }
        g.ncell = (int)ncell; g.periodic = 1; g.z_ring = 1; g.kz_off = 0; g.ncz_global = g.nc[2]; g.row_l0 = 0; g.row_l1 = g.nc[2];
        g.sub_bits = 0;
        ncell_cap = (size_t)ncell;
    } else {
        ncell_cap = std::max<size_t>(4096, std::min<size_t>((size_t)n, (size_t)1 << 22));
        float bb[8];
        shim_launch(1, 32, [&] { bbox_init_kernel(bb); });
        shim_launch(std::min<unsigned>((n + 255) / 256, 1184u), 256, [&] { bbox_kernel(x.data(), n, bb); });
        shim_launch(1, 32, [&] { grid_from_bbox_kernel(bb, (float)cw_min, (int)ncell_cap, &g); });
        g.sub_bits = 0;  // the harness sorts on the plain cell id
    }
    std::vector<uint32_t> keys(n), vals(n), cell_start(ncell_cap + 2, 0);
    shim_launch((n + 255) / 256, 256, [&] { wrap_key_kernel(x.data(), n, &g, keys.data(), vals.data()); });
    // the device radix sort, host driver included, with the key width engine.cu's setup_grid derives
    int bits = 1;
    while (((size_t)1 << bits) < ncell_cap) ++bits;
    if (!periodic) bits += MC_SUB_BITS;
    std::vector<uint32_t> keys1(n), vals1(n), scratch(std::max(radix_scratch_elems((size_t)n), scan_scratch_elems((size_t)n + 1)) + 64);
    uint32_t *kk[2] = {keys.data(), keys1.data()}, *vv[2] = {vals.data(), vals1.data()};
    const int which = radix_sort_pairs(kk, vv, (size_t)n, bits, scratch.data(), nullptr, nullptr);
    std::vector<uint32_t> &skeys = which ? keys1 : keys, &svals = which ? vals1 : vals;
    ReorderArrays ra;
    ra.xyzq_in = x.data(); ra.xyzq_out = xo.data(); ra.xref = xref.data();
    ra.vel_in = vin.data(); ra.vel_out = vout.data();
    ra.type_in = tin.data(); ra.type_out = tout.data();
    ra.flags_in = fin.data(); ra.flags_out = flags_out;
    ra.orig_in = oin.data(); ra.orig_out = orig_out; ra.slot_of_orig = slot_of_orig.data();
    ra.cell_start = cell_start.data();
    ra.mark_interior = 1;
    shim_launch((n + 1 + 255) / 256, 256, [&] { reorder_kernel(n, skeys.data(), svals.data(), &g, ra); });
    const float rl2 = r_list * r_list;
    if (use_tile) {
        // engine_build_rows: the single-pass build with tile and list capacities that adapt on demand.
        // tile_stats: [0] launches, [1] final tile capacity, [2] largest neighbourhood seen, [3] final list capacity
        uint32_t ctl[8];
        uint32_t tile_cap = (uint32_t)tile_cap0;
        std::vector<uint32_t> list((size_t)list_cap0);
        memset(nbr_count, 0, sizeof(uint32_t) * (size_t)n);
        int64_t launches = 0;
        size_t total = 0;
        for (;;) {
            launch_tile_build(n, periodic ? g.ncell : (int)ncell_cap, split, n_sms, xo.data(), cell_start.data(), &g, rl2, rc_inner * rc_inner,
                              orig_out, excl_start, excl_idx, nbr_count, nbr_start, list.data(), false, true, (uint32_t)list.size(), tile_cap, ctl, nullptr,
                              &launches);
            if (ctl[3] != 0) {
                uint32_t need = (ctl[2] + ctl[2] / 4 + 127u) & ~31u;
                if (need > tile_sweep_max_atoms() && ((ctl[2] + 127u) & ~31u) <= tile_sweep_max_atoms()) need = tile_sweep_max_atoms();
                if (need <= tile_sweep_max_atoms()) { tile_cap = need; continue; }
                return -2;  // too dense for the tile path
            }
            total = ctl[1];
            if (total > list.size()) { list.assign(total + total / 8 + 1024, 0u); continue; }
            break;
        }
        tile_stats[0] = (int)launches; tile_stats[1] = (int)tile_cap; tile_stats[2] = (int)ctl[2]; tile_stats[3] = (int)list.size();
        if ((long)total > cap) return -1;
        memcpy(nbr_list, list.data(), sizeof(uint32_t) * total);
        memcpy(xyzq_sorted, xo.data(), sizeof(float4) * (size_t)n);
        for (int a = 0; a < 3; ++a) n_cells_out[a] = g.nc[a];
        return (long)total;
    }
    const unsigned blocks = (unsigned)(((size_t)n * 32 + 255) / 256);
    shim_launch(blocks, 256, [&] {
        sweep_kernel<false>(n, xo.data(), cell_start.data(), &g, rl2, skeys.data(), orig_out, excl_start, excl_idx, nbr_count, nullptr, nullptr);
    });
    // row starts: the device scan with rows padded to 8 entries (nbr_start has n + 1 slots; the total lands in the last)
    std::vector<uint32_t> starts(n + 1);
    exclusive_scan_u32(nbr_count, starts.data(), (size_t)n, 1, scratch.data(), nullptr, nullptr);
    memcpy(nbr_start, starts.data(), sizeof(uint32_t) * (size_t)n);
    const uint64_t total = starts[n];
    if ((long)total > cap) return -1;
    shim_launch(blocks, 256, [&] {
        sweep_kernel<true>(n, xo.data(), cell_start.data(), &g, rl2, skeys.data(), orig_out, excl_start, excl_idx, nbr_count, nbr_start, nbr_list);
    });
    memcpy(xyzq_sorted, xo.data(), sizeof(float4) * (size_t)n);
    for (int a = 0; a < 3; ++a) n_cells_out[a] = g.nc[a];
    return (long)total;
}
}
