// neighbor.cu -- cell list + Verlet neighbour-list build on the device (SURVEY 8a row a3;
// the handle the reference exposes is md.rebuild_spatial_caches(dev),
// properties/sol_shrinking_box.rs:632).
//
// Pipeline (all on the handle's stream, no host round trip for a periodic box):
//   wrap_key      positions wrapped into the box, cell id per atom              16N r + 8N w
//   radix sort    stable (cell id, slot) sort, sort_scan.cu                     2 passes x ~24N
//   reorder       every per-atom array gathered into cell order, cell starts    ~60N
//   count         27-cell sweep, bit-exact accept test, row lengths             L2-resident candidates
//   scan          row starts (rows padded to 8 entries = 32 B sectors)
//   fill          same sweep, ballot-compacted in-order writes                  4 * P_full w
//
// The accept test is the oracle's fp32 expression bit for bit (oracle/md_oracle.c dist2_f32):
//   d -= rintf(d / ext) * ext          (reference src/cuda/util.cu:65-71)
//   r2 = ((dx*dx) + (dy*dy)) + (dz*dz) with no fma contraction; accept r2 < r_list^2, j != i.
#include "common.cuh"
#include "neighbor.cuh"

namespace {

// exact d - rintf(d/ext)*ext without paying for an IEEE division on every candidate: the
// reciprocal estimate decides unless the quotient sits within 1e-4 of a rounding boundary.
__device__ __forceinline__ float min_image_exact(float d, float ext, float inv_ext) {
    float q = __fmul_rn(d, inv_ext);
    float n = rintf(q);
    if (fabsf(q - n) > 0.4999f) n = rintf(__fdiv_rn(d, ext));
    return __fmaf_rn(-n, ext, d);  // n*ext is exact for |n| <= 2, so this is d - n*ext rounded once
}

__device__ __forceinline__ float dist2_exact(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---- bounding box (vacuum systems) ------------------------------------------------------------
__global__ void bbox_kernel(const float4 *__restrict__ xyzq, int n, float *__restrict__ bb /* 6: min xyz, max xyz */) {
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = xyzq[i];
        mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
        mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(MC_FULL_MASK, mn[a], d));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(MC_FULL_MASK, mx[a], d));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        // float atomic min/max through the ordered-int trick (values are finite)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            int *pmn = reinterpret_cast<int *>(bb + a), *pmx = reinterpret_cast<int *>(bb + 3 + a);
            if (mn[a] >= 0.f) atomicMin(pmn, __float_as_int(mn[a])); else atomicMax(reinterpret_cast<unsigned *>(pmn), __float_as_uint(mn[a]));
            if (mx[a] >= 0.f) atomicMax(pmx, __float_as_int(mx[a])); else atomicMin(reinterpret_cast<unsigned *>(pmx), __float_as_uint(mx[a]));
        }
    }
}

__global__ void bbox_init_kernel(float *bb) {
    if (threadIdx.x < 3) bb[threadIdx.x] = 3.0e38f;
    else if (threadIdx.x < 6) bb[threadIdx.x] = -3.0e38f;
}

// vacuum: derive the cell grid from the bounding box; cells never narrower than cw_min, at
// most max_cells in total
__global__ void grid_from_bbox_kernel(const float *__restrict__ bb, float cw_min, int max_cells, GridParams *g) {
    if (threadIdx.x != 0) return;
    float ext[3];
    for (int a = 0; a < 3; ++a) {
        g->lo[a] = bb[a] - 0.5f;
        ext[a] = bb[3 + a] - bb[a] + 1.0f;
        if (!(ext[a] > 0.f && ext[a] < 1.0e7f)) {  // non-finite or absurd coordinates: keep the grid sane,
            g->lo[a] = 0.f;                          // the step loop reports the blow-up (integrate.cu)
            ext[a] = 1.0e7f;
        }
    }
    int nc[3];
    for (int a = 0; a < 3; ++a) {
        nc[a] = (int)floorf(ext[a] / cw_min);
        if (nc[a] < 1) nc[a] = 1;
    }
    while ((long long)nc[0] * nc[1] * nc[2] > max_cells) {  // coarsen the longest axis
        int a = nc[0] >= nc[1] ? (nc[0] >= nc[2] ? 0 : 2) : (nc[1] >= nc[2] ? 1 : 2);
        nc[a] = (nc[a] + 1) / 2;
    }
    for (int a = 0; a < 3; ++a) {
        g->ext[a] = ext[a];
        g->inv_ext[a] = 1.0f / ext[a];
        g->nc[a] = nc[a];
        g->inv_cw[a] = (float)nc[a] / ext[a];
    }
    g->ncell = nc[0] * nc[1] * nc[2];
    g->periodic = 0;
    g->z_ring = 0;
    g->kz_off = 0;
    g->ncz_global = nc[2];
    g->row_l0 = 0;
    g->row_l1 = nc[2];
    g->sub_bits = MC_SUB_BITS;  // max_cells <= 2^22: the key stays within 32 bits
}

// ---- wrap + cell key --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wrap_key_kernel(float4 *__restrict__ xyzq, int n,
                                                        const GridParams *__restrict__ gp,
                                                        uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const GridParams g = *gp;
    float4 p = xyzq[i];
    float c[3] = {p.x, p.y, p.z};
    float fr[3];
    int k[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (g.periodic) {
            c[a] -= floorf((c[a] - g.lo[a]) * g.inv_ext[a]) * g.ext[a];
            // one corrective step for the rounding cases of the line above
            if (c[a] < g.lo[a]) c[a] += g.ext[a];
            if (c[a] >= g.lo[a] + g.ext[a]) c[a] -= g.ext[a];
        }
        const float u = (c[a] - g.lo[a]) * g.inv_cw[a];
        int kk = (int)floorf(u);
        k[a] = min(max(kk, 0), g.nc[a] - 1);
        fr[a] = u - (float)k[a];
    }
    if (g.periodic) xyzq[i] = make_float4(c[0], c[1], c[2], p.w);
    const uint32_t cell = (uint32_t)((k[2] * g.nc[1] + k[1]) * g.nc[0] + k[0]);
    keys[i] = g.sub_bits ? (cell << MC_SUB_BITS) | mc_subcell_code(fr[0], fr[1], fr[2]) : cell;
    vals[i] = (uint32_t)i;
}

// ---- reorder into cell order + cell starts ------------------------------------------------------
__global__ void __launch_bounds__(256) reorder_kernel(int n, const uint32_t *__restrict__ skeys,
                                                       const uint32_t *__restrict__ svals,
                                                       const GridParams *__restrict__ gp, ReorderArrays a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n) return;
    const int ncell = gp->ncell, sb = gp->sub_bits;
    // cell_start[c] = first sorted slot whose cell >= c  (boundary detection on the sorted keys)
    const int prev = k == 0 ? -1 : (int)(skeys[k - 1] >> sb);
    const int cur = k == n ? ncell : (int)(skeys[k] >> sb);
    for (int c = prev + 1; c <= cur; ++c) a.cell_start[c] = (uint32_t)k;
    if (k == n) return;
    const uint32_t src = svals[k];
    const float4 p = a.xyzq_in[src];
    a.xyzq_out[k] = p;
    a.xref[k] = p;
    a.vel_out[k] = a.vel_in[src];
    a.type_out[k] = a.type_in[src];
    // MC_FLAG_INTERIOR: the cell's 27-cell stencil never wraps around the box, so every listed
    // partner stays within ext/2 of this atom until the next build and the force kernel can skip
    // the minimum image for this row (d - rintf(d/ext)*ext with n == 0 is d itself, bit for bit).
    uint8_t fl = a.flags_in[src] & (uint8_t)~MC_FLAG_INTERIOR;
    if (gp->periodic && a.mark_interior) {
        const int nc0 = gp->nc[0], nc1 = gp->nc[1], nc2 = gp->ncz_global;
        const int c0 = cur % nc0, c1 = (cur / nc0) % nc1, c2 = (cur / (nc0 * nc1) + gp->kz_off) % nc2;  // global layer
        if (nc0 >= 3 && nc1 >= 3 && nc2 >= 3 && c0 >= 1 && c0 <= nc0 - 2 && c1 >= 1 && c1 <= nc1 - 2 && c2 >= 1 &&
            c2 <= nc2 - 2)
            fl |= MC_FLAG_INTERIOR;
    }
    a.flags_out[k] = fl;
    const int o = a.orig_in[src];
    a.orig_out[k] = o;
    a.slot_of_orig[o] = k;
}

// ---- neighbour sweep ----------------------------------------------------------------------------
// One warp per atom i (sorted slot).  Candidate cells of one (dz, dy) row are contiguous in the
// sorted order, so the 27 cells are visited as <= 9 (+ wrap splits) contiguous slot ranges and
// the 32 lanes read consecutive float4 -- 512-byte coalesced requests that hit L1/L2 because the
// neighbouring warps sweep the same ranges.  FILL == false counts, FILL == true writes the row
// with a ballot compaction that preserves the sweep order.
template <bool FILL>
__global__ void __launch_bounds__(256) sweep_kernel(int n_rows, const float4 *__restrict__ xyzq,
                                                     const uint32_t *__restrict__ cell_start,
                                                     const GridParams *__restrict__ gp, float rl2,
                                                     const uint32_t *__restrict__ cell_of_slot,
                                                     const int *__restrict__ orig,
                                                     const int32_t *__restrict__ excl_start,
                                                     const int32_t *__restrict__ excl_idx,
                                                     uint32_t *__restrict__ nbr_count,
                                                     const uint32_t *__restrict__ nbr_start,
                                                     uint32_t *__restrict__ nbr_list) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_rows) return;
    const GridParams g = *gp;
    const float4 pi = xyzq[i];
    // the (local) cell of this slot comes from the sorted keys: on a decomposed rank the local z layer
    // is not a function of the coordinate alone
    const int cell = (int)(cell_of_slot[i] >> g.sub_bits);
    int ci[3] = {cell % g.nc[0], (cell / g.nc[0]) % g.nc[1], cell / (g.nc[0] * g.nc[1])};
    if (ci[2] < g.row_l0 || ci[2] >= g.row_l1) {  // ghost layer: no row
        if (!FILL && lane == 0) nbr_count[i] = 0;
        return;
    }
    int ex_lo = 0, ex_hi = 0;
    if (excl_start) {
        const int oi = orig[i];
        ex_lo = excl_start[oi];
        ex_hi = excl_start[oi + 1];
    }
    uint32_t cnt = 0;
    const uint32_t row = FILL ? nbr_start[i] : 0u;

    // offsets per axis, de-duplicated for tiny periodic grids (oracle/md_oracle.c does the same)
    const int lo_y = (g.nc[1] >= 3 || !g.periodic) ? -1 : 0, hi_y = (g.nc[1] >= 2 || !g.periodic) ? 1 : 0;
    const int lo_z = (g.nc[2] >= 3 || !g.z_ring) ? -1 : 0, hi_z = (g.nc[2] >= 2 || !g.z_ring) ? 1 : 0;

    for (int dz = lo_z; dz <= hi_z; ++dz) {
        int kz = ci[2] + dz;
        if (g.z_ring) kz = (kz + g.nc[2]) % g.nc[2];
        else if (kz < 0 || kz >= g.nc[2]) continue;
        for (int dy = lo_y; dy <= hi_y; ++dy) {
            int ky = ci[1] + dy;
            if (g.periodic) ky = (ky + g.nc[1]) % g.nc[1];
            else if (ky < 0 || ky >= g.nc[1]) continue;
            const int rowbase = (kz * g.nc[1] + ky) * g.nc[0];
            // x cells: contiguous run [x0, x1] plus (periodic only) a wrapped remainder
            int runs[2][2];
            int nruns = 1;
            if (!g.periodic || g.nc[0] < 3) {
                if (g.periodic) { runs[0][0] = 0; runs[0][1] = g.nc[0] - 1; }  // 1 or 2 cells: all of them
                else { runs[0][0] = max(ci[0] - 1, 0); runs[0][1] = min(ci[0] + 1, g.nc[0] - 1); }
            } else if (ci[0] == 0) {
                runs[0][0] = 0; runs[0][1] = 1; runs[1][0] = runs[1][1] = g.nc[0] - 1; nruns = 2;
            } else if (ci[0] == g.nc[0] - 1) {
                runs[0][0] = 0; runs[0][1] = 0; runs[1][0] = g.nc[0] - 2; runs[1][1] = g.nc[0] - 1; nruns = 2;
            } else {
                runs[0][0] = ci[0] - 1; runs[0][1] = ci[0] + 1;
            }
            for (int r = 0; r < nruns; ++r) {
                const uint32_t s0 = cell_start[rowbase + runs[r][0]];
                const uint32_t s1 = cell_start[rowbase + runs[r][1] + 1];
                for (uint32_t s = s0; s < s1; s += 32) {
                    const uint32_t j = s + lane;
                    bool hit = false;
                    if (j < s1 && j != (uint32_t)i) {
                        const float4 pj = xyzq[j];
                        float dx = __fsub_rn(pi.x, pj.x), dy_ = __fsub_rn(pi.y, pj.y), dz_ = __fsub_rn(pi.z, pj.z);
                        if (g.periodic) {
                            dx = min_image_exact(dx, g.ext[0], g.inv_ext[0]);
                            dy_ = min_image_exact(dy_, g.ext[1], g.inv_ext[1]);
                            dz_ = min_image_exact(dz_, g.ext[2], g.inv_ext[2]);
                        }
                        hit = dist2_exact(dx, dy_, dz_) < rl2;
                        if (hit && ex_hi > ex_lo) {
                            const int oj = orig[j];
                            for (int e = ex_lo; e < ex_hi; ++e)
                                if (excl_idx[e] == oj) { hit = false; break; }
                        }
                    }
                    const uint32_t m = __ballot_sync(MC_FULL_MASK, hit);
                    if (FILL && hit) nbr_list[row + cnt + __popc(m & ((1u << lane) - 1u))] = j;
                    cnt += __popc(m);
                }
            }
        }
    }
    if (!FILL && lane == 0) nbr_count[i] = cnt;
}

// ---- export (mc_get_neighbors): rows back in original ids -----------------------------------------
__global__ void count_by_orig_kernel(int n, const int *__restrict__ orig, const uint32_t *__restrict__ nbr_count,
                                     uint32_t *__restrict__ cnt_orig) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) cnt_orig[orig[k]] = nbr_count[k];
}

__global__ void translate_rows_kernel(int n, const int *__restrict__ orig, const uint32_t *__restrict__ nbr_count,
                                      const uint32_t *__restrict__ nbr_start, const uint32_t *__restrict__ nbr_list,
                                      const uint32_t *__restrict__ start_orig, uint32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n) return;
    const uint32_t c = nbr_count[k], src = nbr_start[k], dst = start_orig[orig[k]];
    for (uint32_t t = lane; t < c; t += 32) out[dst + t] = (uint32_t)orig[nbr_list[src + t]];
}

// in-place ascending sort of every row: one warp per row, bitonic passes through shared memory for
// rows up to ROW_SMEM entries, odd-even transposition in global memory beyond that (export path
// only, never on the step path)
constexpr int ROW_SMEM = 2048;
__global__ void __launch_bounds__(128) sort_rows_kernel(int n, const uint32_t *__restrict__ start, uint32_t *__restrict__ rows) {
    __shared__ uint32_t buf[4][ROW_SMEM];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r = blockIdx.x * 4 + w;
    if (r >= n) return;
    const uint32_t s = start[r], c = start[r + 1] - s;
    if (c <= 1) return;
    if (c <= ROW_SMEM) {
        uint32_t m = 1;
        while (m < c) m <<= 1;
        for (uint32_t t = lane; t < m; t += 32) buf[w][t] = t < c ? rows[s + t] : 0xffffffffu;
        __syncwarp();
        for (uint32_t k = 2; k <= m; k <<= 1) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t t = lane; t < m; t += 32) {
                    uint32_t p = t ^ j;
                    if (p > t) {
                        uint32_t a = buf[w][t], b = buf[w][p];
                        bool up = (t & k) == 0;
                        if ((a > b) == up) { buf[w][t] = b; buf[w][p] = a; }
                    }
                }
                __syncwarp();
            }
        }
        for (uint32_t t = lane; t < c; t += 32) rows[s + t] = buf[w][t];
    } else {
        for (uint32_t pass = 0; pass < c; ++pass) {
            for (uint32_t t = (pass & 1) + 2 * lane; t + 1 < c; t += 64) {
                uint32_t a = rows[s + t], b = rows[s + t + 1];
                if (a > b) { rows[s + t] = b; rows[s + t + 1] = a; }
            }
            __syncwarp();
        }
    }
}

}  // namespace

#ifdef MC_HAVE_LAUNCH  // the stand-ins of tests/cpp/shim/ and shim_mt/ have no launcher; shim_fiber/ has
void launch_bbox(const float4 *xyzq, int n, float *bb, float cw_min, int max_cells, GridParams *g, cudaStream_t st,
                 int64_t *launches) {
    MC_LAUNCH(bbox_init_kernel, 1, 32, 0, st, bb);
    MC_LAUNCH(bbox_kernel, min(div_up(n, 256), 1184u), 256, 0, st, xyzq, n, bb);
    MC_LAUNCH(grid_from_bbox_kernel, 1, 32, 0, st, bb, cw_min, max_cells, g);
    *launches += 3;
}

void launch_wrap_key(float4 *xyzq, int n, const GridParams *g, uint32_t *keys, uint32_t *vals, cudaStream_t st,
                     int64_t *launches) {
    MC_LAUNCH(wrap_key_kernel, div_up(n, 256), 256, 0, st, xyzq, n, g, keys, vals);
    *launches += 1;
}

void launch_reorder(int n, const uint32_t *skeys, const uint32_t *svals, const GridParams *g, const ReorderArrays &a,
                    cudaStream_t st, int64_t *launches) {
    MC_LAUNCH(reorder_kernel, div_up((size_t)n + 1, 256), 256, 0, st, n, skeys, svals, g, a);
    *launches += 1;
}

void launch_sweep(bool fill, int n_rows, const float4 *xyzq, const uint32_t *cell_start, const GridParams *g, float rl2,
                  const uint32_t *cell_of_slot, const int *orig, const int32_t *excl_start, const int32_t *excl_idx, uint32_t *nbr_count,
                  const uint32_t *nbr_start, uint32_t *nbr_list, cudaStream_t st, int64_t *launches) {
    const unsigned blocks = div_up((size_t)n_rows * 32, 256);
    if (fill)
        MC_LAUNCH(sweep_kernel<true>, blocks, 256, 0, st, n_rows, xyzq, cell_start, g, rl2, cell_of_slot, orig, excl_start, excl_idx, nbr_count,
                                                   nbr_start, nbr_list);
    else
        MC_LAUNCH(sweep_kernel<false>, blocks, 256, 0, st, n_rows, xyzq, cell_start, g, rl2, cell_of_slot, orig, excl_start, excl_idx, nbr_count,
                                                    nbr_start, nbr_list);
    *launches += 1;
}

void launch_export_rows(int n, const int *orig, const uint32_t *nbr_count, const uint32_t *nbr_start,
                        const uint32_t *nbr_list, uint32_t *cnt_orig, uint32_t *start_orig, uint32_t *rows,
                        uint32_t *scan_scratch, cudaStream_t st, int64_t *launches) {
    MC_LAUNCH(count_by_orig_kernel, div_up(n, 256), 256, 0, st, n, orig, nbr_count, cnt_orig);
    exclusive_scan_u32(cnt_orig, start_orig, n, 0, scan_scratch, st, launches);
    MC_LAUNCH(translate_rows_kernel, div_up((size_t)n * 32, 256), 256, 0, st, n, orig, nbr_count, nbr_start, nbr_list, start_orig, rows);
    MC_LAUNCH(sort_rows_kernel, div_up(n, 4), 128, 0, st, n, start_orig, rows);
    *launches += 3;
}
#endif  // MC_HOST_SHIM
