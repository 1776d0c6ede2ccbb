/**
 *  @file nccl.h (tests/emulator)
 *  @brief TEST INFRASTRUCTURE: the NCCL types the host driver names. The emulated library resolves NCCL at run time like
 *         the real one (dlopen), and the CPU test tier never creates a communicator.
 */
#pragma once
#include <cuda_runtime.h>

typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1, ncclSystemError = 2, ncclInternalError = 3 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1 } ncclDataType_t;
