/* Test infrastructure: the slice of <nccl.h> that molchanica_b200/csrc/nccl_dyn.cuh and comm.cu use, for the host build of
 * the library (tests/cpp/host_lib/).  The entry points themselves come from tests/cpp/host_lib/nccl_standin.cpp through the
 * same dlopen as the real library (MOLCHANICA_NCCL_LIB). */
#pragma once
#ifndef MC_NCCL_STANDIN_NO_RUNTIME
#include <cuda_runtime.h>
#endif
#include <stddef.h>
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1, ncclSystemError = 2, ncclInternalError = 3, ncclInvalidArgument = 4 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclInt = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5,
               ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat = 7, ncclFloat64 = 8, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
