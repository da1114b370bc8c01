// common.cuh -- shared declarations of the sm_100a engine behind include/molchanica_md.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/molchanica_md.h"

// Dynamic shared memory of a kernel.  The host stand-ins of tests/cpp/ (MC_HOST_SHIM) map it onto a process-wide buffer.
#ifndef MC_HOST_SHIM
#define MC_DYN_SHARED(T, name) extern __shared__ T name[]
#define MC_DYN_SHARED_ALIGNED(T, name, A) extern __shared__ __align__(A) T name[]
#else
#define MC_DYN_SHARED(T, name) T *name = reinterpret_cast<T *>(shim_dyn_smem)
#define MC_DYN_SHARED_ALIGNED(T, name, A) T *name = reinterpret_cast<T *>(shim_dyn_smem)
#endif

// MC_HOST_SHIM: the file is being compiled by g++ over one of the cuda_runtime.h stand-ins of tests/cpp/ (no PTX).
// MC_HOST_LAUNCH: that stand-in can also run launches and the host runtime (tests/cpp/shim_fiber/), so the launchers and
// engine.cu are compiled as well.
#if !defined(MC_HOST_SHIM) || defined(MC_HOST_LAUNCH)
#define MC_HAVE_LAUNCH 1
#endif

// Kernel launch.  Host drivers written with it run unchanged over the stand-ins (one block at a time on OS threads).
#ifndef MC_HOST_SHIM
#define MC_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#else
#define MC_LAUNCH(kernel, grid, block, smem, stream, ...) shim_launch((grid), (block), [&] { kernel(__VA_ARGS__); })
#endif

#define MC_COMMA ,  // for kernel names with several template arguments inside MC_LAUNCH

#define MC_WARP 32
#define MC_FULL_MASK 0xffffffffu
#define MC_ACCEL_CONV 418.4f       // kcal/mol/A/amu -> A/ps^2
#define MC_SOFTENING_SQ 0.000001f  // reference src/cuda/util.cu:9-10
#define MC_INV_SQRT_PI 0.5641895835477563f  // util.cu:15-18
#define MC_KB 0.0019872041         // kcal/mol/K
#define MC_BAR_PER_KCAL_MOL_A3 69476.95  // 1 kcal/mol/A^3 in bar
#define MC_FLAG_INTERIOR 0x80u      // engine-internal flag bit: atom sits in a cell whose 27-cell stencil never wraps

// Device-resident description of the cell grid; written by the host (periodic box) or by
// grid_from_bbox_kernel (vacuum), read by every neighbour-build kernel.
struct GridParams {
    float lo[3];       // origin
    float ext[3];      // box extent (periodic) / bounding extent (vacuum)
    float inv_ext[3];
    float inv_cw[3];   // 1 / cell width
    int nc[3];         // local grid; nc[2] counts the LOCAL z layers
    int ncell;
    int periodic;      // minimum image in the pair distance, stencil wraps in x and y
    int z_ring;        // 1: the local z layers are the whole periodic ring (stencil wraps in z);
                       // 0: a segment of it (domain decomposition: owned layers + one ghost layer each side) or vacuum
    int kz_off;        // global z layer of local layer 0
    int ncz_global;    // global number of z layers (== nc[2] unless decomposed)
    int row_l0, row_l1;  // local z layers [row_l0, row_l1) carry list rows (the others are ghost layers)
    int sub_bits;      // low bits of the sort key: Morton code of the atom's sub-cell (4 x 4 x 4 per cell), so that
                       // atoms that are close in space are close in memory and a row's gathers share cache lines
};

#define MC_SUB_BITS 6

// Morton code (2 bits per axis) of the fractional position f in [0,1)^3 inside a cell.
__host__ __device__ inline uint32_t mc_subcell_code(float fx, float fy, float fz) {
    const int sx = fx < 0.f ? 0 : (fx >= 1.f ? 3 : (int)(fx * 4.f));
    const int sy = fy < 0.f ? 0 : (fy >= 1.f ? 3 : (int)(fy * 4.f));
    const int sz = fz < 0.f ? 0 : (fz >= 1.f ? 3 : (int)(fz * 4.f));
    auto part = [](int v) { return (uint32_t)(((v & 2) << 2) | (v & 1)); };
    return part(sx) | (part(sy) << 1) | (part(sz) << 2);
}

// Nonbonded parameters passed by value to the pair kernels.
struct NbParams {
    float ext[3], inv_ext[3];
    float rc2_lj, rc2_q;   // squared cutoffs (fp32 products, same rounding as the oracle)
    float alpha;           // Ewald splitting parameter
    float sig2, eps24;     // single-type fast path: sigma^2, 24*eps
    int periodic;
    int n_types;
};

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ---- sort_scan.cu -------------------------------------------------------------------------
// Exclusive prefix sum of f(in[i]) into out[0..n] (out[n] = total).  align8 != 0 rounds every
// input up to a multiple of 8 first (row starts of the Verlet list are 32-byte aligned).
// scratch: >= scan_scratch_elems(n) uint32.
size_t scan_scratch_elems(size_t n);
void exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n, int align8, uint32_t *scratch,
                        cudaStream_t st, int64_t *launches);
// Stable LSD radix sort of (key, val) pairs on `bits` low key bits.  Results end in keys[0] /
// vals[0] when the returned value is 0, in keys[1] / vals[1] when it is 1.
size_t radix_scratch_elems(size_t n);
int radix_sort_pairs(uint32_t *keys[2], uint32_t *vals[2], size_t n, int bits, uint32_t *scratch,
                     cudaStream_t st, int64_t *launches);
