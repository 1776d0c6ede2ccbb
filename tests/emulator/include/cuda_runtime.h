/**
 *  @file cuda_runtime.h (tests/emulator)
 *  @brief TEST INFRASTRUCTURE, never part of the library: the slice of the CUDA runtime API and of the device-side
 *         language that `stringzilla_b200/csrc/*` uses, re-declared so that the UNCHANGED kernels and host driver compile
 *         with g++ and run on the CPU test tier under a lock-step SIMT emulator (simt.cpp): one fiber per CUDA thread,
 *         warp collectives and CTA barriers as rendezvous points, device memory as plain host memory with a registry
 *         that answers `cudaPointerGetAttributes`.
 *
 *  Built by tests/emulator/build.py into tests/emulator/_build/libstringzillas_b200_emulated.so; only tests load it.
 *  About 10^4 times slower than a B200: a parity checker for kernel LOGIC, not a fallback.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* ---- Language keywords ------------------------------------------------------------------------------------------- */
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
/** Static shared arrays become function-local statics: one CTA runs at a time, and every kernel initialises them. */
#define __shared__ static

/* ---- Runtime API: types ------------------------------------------------------------------------------------------ */
enum cudaError_t {
    cudaSuccess = 0,
    cudaErrorInvalidValue = 1,
    cudaErrorMemoryAllocation = 2,
    cudaErrorInsufficientDriver = 35,
    cudaErrorNoDevice = 100,
    cudaErrorInvalidDevice = 101,
};
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrComputeCapabilityMajor = 75, cudaDevAttrComputeCapabilityMinor = 76 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
struct cudaPointerAttributes {
    cudaMemoryType type;
    int device;
};
typedef struct simt_stream *cudaStream_t;
typedef struct simt_event *cudaEvent_t;
struct uint4 {
    unsigned x, y, z, w;
};
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4 {x, y, z, w}; }

/* ---- Runtime API: functions (simt.cpp). Copies and kernels complete before the call returns: a sequentially consistent
 *      execution, which is one of the executions CUDA's stream semantics allow. ------------------------------------------ */
cudaError_t cudaGetLastError();
char const *cudaGetErrorName(cudaError_t);
cudaError_t cudaGetDeviceCount(int *);
cudaError_t cudaSetDevice(int);
cudaError_t cudaGetDevice(int *);
cudaError_t cudaDeviceGetAttribute(int *, cudaDeviceAttr, int device);
cudaError_t cudaMalloc(void **, size_t);
cudaError_t cudaMallocHost(void **, size_t);
cudaError_t cudaMallocManaged(void **, size_t);
cudaError_t cudaFree(void *);
cudaError_t cudaFreeHost(void *);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *, void const *);
cudaError_t cudaMemcpy(void *, void const *, size_t, cudaMemcpyKind);
cudaError_t cudaMemcpyAsync(void *, void const *, size_t, cudaMemcpyKind, cudaStream_t);
cudaError_t cudaMemcpy2DAsync(void *, size_t, void const *, size_t, size_t width, size_t height, cudaMemcpyKind, cudaStream_t);
cudaError_t cudaMemsetAsync(void *, int, size_t, cudaStream_t);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *, unsigned);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned);
cudaError_t cudaEventCreate(cudaEvent_t *);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t);
cudaError_t cudaEventElapsedTime(float *, cudaEvent_t, cudaEvent_t);
template <typename kernel_type_>
cudaError_t cudaFuncSetAttribute(kernel_type_, cudaFuncAttribute, int) {
    return cudaSuccess;
}

/* ---- The SIMT machine (simt.cpp) ----------------------------------------------------------------------------------- */
namespace simt {

struct Index3 {
    unsigned x, y, z;
};
/** Identity of the CUDA thread whose fiber is running. */
extern Index3 current_thread, current_block, current_block_dim, current_grid_dim;
/** The running CTA's dynamic shared memory (227 KB, 128-byte aligned). */
uint32_t *dynamic_shared();

/** Runs `grid` CTAs of `threads` fibers one CTA after the other; `body(context)` is the kernel body of ONE thread. */
void launch(unsigned grid, unsigned threads, size_t shared_bytes, void (*body)(void *), void *context);
template <typename body_type_>
void launch(unsigned grid, unsigned threads, size_t shared_bytes, body_type_ &&body) {
    launch(grid, threads, shared_bytes, [](void *context) { (*static_cast<body_type_ *>(context))(); }, (void *)&body);
}

void barrier_cta();
/** Blocks the calling thread while `*word == value` (another thread of the CTA changes it). */
void wait_while_equal(uint32_t const volatile *word, uint32_t value);
/** Warp rendezvous: every live lane deposits `value`; returns a pointer to the 32 deposits of this collective (lanes that
 *  left the kernel read as zero). `kind` tags the call site: lanes of one warp meeting in different collectives is a bug. */
uint32_t const *exchange(uint32_t value, int kind);

} // namespace simt

#define threadIdx (simt::current_thread)
#define blockIdx (simt::current_block)
#define blockDim (simt::current_block_dim)
#define gridDim (simt::current_grid_dim)

/* ---- Device-side builtins ------------------------------------------------------------------------------------------ */
static inline void __syncthreads() { simt::barrier_cta(); }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { simt::exchange(0, 1); }

template <typename value_type_>
static inline uint32_t simt_bits(value_type_ value) {
    static_assert(sizeof(value_type_) == 4, "32-bit shuffles only");
    uint32_t bits;
    memcpy(&bits, &value, 4);
    return bits;
}
template <typename value_type_>
static inline value_type_ simt_value(uint32_t bits) {
    value_type_ value;
    memcpy(&value, &bits, 4);
    return value;
}
template <typename value_type_>
static inline value_type_ __shfl_sync(unsigned, value_type_ value, int source_lane, int width = 32) {
    int const lane = (int)(threadIdx.x & 31), base = lane & ~(width - 1);
    return simt_value<value_type_>(simt::exchange(simt_bits(value), 2)[base + (source_lane & (width - 1))]);
}
template <typename value_type_>
static inline value_type_ __shfl_up_sync(unsigned, value_type_ value, unsigned delta, int width = 32) {
    int const lane = (int)(threadIdx.x & 31), base = lane & ~(width - 1);
    uint32_t const *const all = simt::exchange(simt_bits(value), 3);
    return lane - (int)delta >= base ? simt_value<value_type_>(all[lane - delta]) : value;
}
template <typename value_type_>
static inline value_type_ __shfl_xor_sync(unsigned, value_type_ value, int lane_mask, int width = 32) {
    int const lane = (int)(threadIdx.x & 31), base = lane & ~(width - 1), other = lane ^ lane_mask;
    uint32_t const *const all = simt::exchange(simt_bits(value), 4);
    return (other & ~(width - 1)) == base ? simt_value<value_type_>(all[other]) : value;
}
static inline unsigned __ballot_sync(unsigned, int predicate) {
    uint32_t const *const all = simt::exchange(predicate ? 1u : 0u, 5);
    unsigned ballot = 0;
    for (int lane = 0; lane < 32; ++lane) ballot |= (all[lane] & 1u) << lane;
    return ballot;
}
static inline unsigned __match_any_sync(unsigned, uint32_t value) {
    uint32_t all[32]; // a copy: the next rendezvous reuses the thread's snapshot
    memcpy(all, simt::exchange(value, 6), sizeof(all));
    uint32_t const *const live = simt::exchange(1u, 7); // lanes that left the kernel match nobody
    unsigned peers = 0;
    for (int lane = 0; lane < 32; ++lane) peers |= (unsigned)(live[lane] && all[lane] == value) << lane;
    return peers;
}

template <typename value_type_>
static inline value_type_ atomicAdd(value_type_ *address, value_type_ value) {
    value_type_ const old = *address;
    *address = old + value;
    return old;
}
template <typename value_type_>
static inline value_type_ atomicOr(value_type_ *address, value_type_ value) {
    value_type_ const old = *address;
    *address = old | value;
    return old;
}
template <typename value_type_>
static inline value_type_ atomicCAS(value_type_ *address, value_type_ compare, value_type_ value) {
    value_type_ const old = *address;
    if (old == compare) *address = value;
    return old;
}

static inline int __popc(unsigned value) { return __builtin_popcount(value); }
static inline int __ffs(int value) { return __builtin_ffs(value); }
static inline unsigned __funnelshift_l(unsigned low, unsigned high, unsigned shift) {
    shift &= 31;
    return shift ? (high << shift) | (low >> (32 - shift)) : high;
}
template <typename value_type_>
static inline value_type_ min(value_type_ a, value_type_ b) {
    return b < a ? b : a;
}
template <typename value_type_>
static inline value_type_ max(value_type_ a, value_type_ b) {
    return a < b ? b : a;
}

/* DPX and SIMD-in-a-word intrinsics: what VIMNMX3 / VIADDMNMX compute, per 16-bit half or on 32 bits. */
static inline unsigned __vmaxu2(unsigned a, unsigned b) {
    unsigned const low = max(a & 0xFFFFu, b & 0xFFFFu), high = max(a >> 16, b >> 16);
    return low | (high << 16);
}
static inline unsigned __vminu2(unsigned a, unsigned b) {
    unsigned const low = min(a & 0xFFFFu, b & 0xFFFFu), high = min(a >> 16, b >> 16);
    return low | (high << 16);
}
static inline unsigned __vimax3_u16x2(unsigned a, unsigned b, unsigned c) { return __vmaxu2(__vmaxu2(a, b), c); }
static inline unsigned __viaddmax_u16x2(unsigned a, unsigned b, unsigned c) {
    unsigned const sum = ((a + b) & 0xFFFFu) | (((a >> 16) + (b >> 16)) << 16); // per-half wrap-around add
    return __vmaxu2(sum, c);
}
static inline int __vimax3_s32(int a, int b, int c) { return max(max(a, b), c); }
static inline int __viaddmax_s32(int a, int b, int c) { return max((int)((unsigned)a + (unsigned)b), c); }
