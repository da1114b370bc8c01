/* Test infrastructure: the slice of <cuda.h> comm.cu names (driver types for cuMemGetAddressRange).  The host stand-in has
 * no peer memory: cudaGetDriverEntryPoint and the cudaIpc calls of cuda_runtime.h fail, comm.cu falls back to the NCCL halo. */
#pragma once
#include <stdint.h>
typedef uintptr_t CUdeviceptr;
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
