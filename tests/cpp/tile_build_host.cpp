// Test infrastructure: tile_build.cu -- the production Verlet-list build (persistent, warp-specialised: a producer warp
// stages each cell's 27-cell neighbourhood by bulk copies that complete on an mbarrier, eight consumer warps sweep the
// staged tile and write partitioned rows; 2-stage ring) -- compiled unchanged for the host.  The five PTX wrappers
// (mbarrier init / arrive / expect_tx / try_wait, cp.async.bulk) and the consumers' named barrier map onto
// tests/cpp/shim_mt/mbarrier.h; everything else (work distribution, range lookup, the exact accept test, exclusions,
// row allocation, the inner / skin-shell partition, tile- and list-overflow reporting, the launcher) is the source the
// GPU runs.  A second translation unit next to neighbor_kernels_host.cpp because the two .cu files define helpers of
// the same name; it exports launch_tile_build / tile_sweep_max_atoms as declared in neighbor.cuh.
// Not part of the product library.
#define MC_HOST_SHIM 1
#define MC_SHIM_SHARED_STATIC 1
#include "shim_mt/cuda_runtime.h"
#include "shim_mt/mbarrier.h"

#include "../../molchanica_b200/csrc/tile_build.cu"
