/**
 *  @file engine.cu
 *  @brief Host side of `libstringzillas_b200.so`: engines, device scopes, input staging, the slot/job scheduler and
 *         the `szs_*` C ABI declared in include/stringzillas_b200.h.
 *
 *  Mirrors the roles (not the code) of the reference's C shims and CUDA engine drivers:
 *    c/stringzillas/{runtime,levenshtein,needleman_wunsch,smith_waterman}.cuh and
 *    include/stringzillas/similarities/cuda.cuh:4247-4417 (`cross_`), 5913-6032 (`cuda_weighted_cross_`).
 *  Where the reference materialises one 112-byte task per matrix cell, sorts tasks through six tiers with host
 *  round-trips and scatters results (cuda.cuh:1653-1711, 1916-1955, 4394-4401), this driver keeps the work list at
 *  O(candidates): candidates are sorted by length and paired into slots, each slot is crossed with blocks of queries,
 *  and persistent CTAs drain that job list through one atomic counter. Results are written in place by the kernel.
 *
 *  There is NO CPU fallback in this file: without a usable sm_100-class device every scoring call fails loudly.
 */
#include "../../include/stringzillas_b200.h"
#include "host_scan.hpp"
#include "launch.cuh"
#include "myers.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>

namespace szb {

#pragma region Helper kernels

/** Maps raw bytes through a 256-entry class table (NW/SW: the engine's `byte_to_class`; Levenshtein: the alphabet). */
__global__ void map_classes_kernel(uint8_t const *in, uint8_t *out, uint64_t n, uint8_t const *lut) {
    __shared__ uint8_t table[256];
    if (threadIdx.x < 256) table[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x * 16;
    for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; base < n; base += stride) {
        if (base + 16 <= n) {
            uint4 v = *(uint4 const *)(in + base); // staging buffers are 256-byte aligned
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                w[i] = (uint32_t)table[w[i] & 0xFF] | ((uint32_t)table[(w[i] >> 8) & 0xFF] << 8) |
                       ((uint32_t)table[(w[i] >> 16) & 0xFF] << 16) | ((uint32_t)table[w[i] >> 24] << 24);
            *(uint4 *)(out + base) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        else {
            for (uint64_t i = base; i < n; ++i) out[i] = table[in[i]];
        }
    }
}

/** Callback sequences whose strings live in managed or device memory: one warp per string copies it into the blob. */
__global__ void gather_strings_kernel(uint8_t const *const *starts, uint32_t const *offsets, uint32_t count, uint8_t *out) {
    uint32_t const lane = threadIdx.x & 31, warps = gridDim.x * (blockDim.x / 32);
    for (uint32_t i = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32; i < count; i += warps) {
        uint32_t const begin = offsets[i], length = offsets[i + 1] - begin;
        uint8_t const *const source = starts[i];
        for (uint32_t b = lane; b < length; b += 32) out[begin + b] = source[b];
    }
}

/** Marks which byte values occur in a blob: flags[b] != 0. Drives the Levenshtein alphabet compaction. */
__global__ void byte_presence_kernel(uint8_t const *in, uint64_t n, uint32_t *flags) {
    __shared__ uint32_t seen[256];
    if (threadIdx.x < 256) seen[threadIdx.x] = 0;
    __syncthreads();
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) seen[in[i]] = 1;
    __syncthreads();
    if (threadIdx.x < 256 && seen[threadIdx.x]) flags[threadIdx.x] = 1;
}

/** Role-swapped calls score (candidate, query): `swapped` is (columns x rows), `home` becomes (rows x columns). */
__global__ void transpose_results_kernel(int64_t const *swapped, int64_t *home, uint64_t rows, uint64_t columns) {
    uint64_t const total = rows * columns;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t const row = idx / columns, column = idx % columns;
        home[idx] = swapped[column * rows + row];
    }
}

/** Symmetric mode: mirror the lower triangle (query i, candidate j < i) into the upper one (serial.hpp:3183-3184). */
__global__ void mirror_lower_kernel(int64_t *results, uint64_t n, uint64_t stride) {
    uint64_t const total = n * n;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t const i = idx / n, j = idx % n;
        if (j > i) results[i * stride + j] = results[j * stride + i];
    }
}

/**
 *  Symmetric mode with symmetric costs: only cells (query q, candidate c) with position[q] <= position[c] in the
 *  longest-first query order were scored (each unordered pair once, the longer string as the query); fill the rest.
 */
__global__ void mirror_ranked_kernel(int64_t *results, uint64_t n, uint64_t stride, uint32_t const *position) {
    uint64_t const total = n * n;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t const i = idx / n, j = idx % n;
        if (position[i] > position[j]) results[i * stride + j] = results[j * stride + i];
    }
}

/**
 *  UTF-8 -> UTF-32, one thread per string, with the reference's UNCHECKED decoding rule
 *  (include/stringzilla/utf8_runes/serial.h:111-124): the lead byte alone decides the sequence length
 *  (1 + (lead >= 0xC0) + (lead >= 0xE0) + (lead >= 0xF0)); bytes past the end of the string read as zero.
 *  `runes == nullptr` only counts.
 */
__global__ void utf8_decode_kernel(uint8_t const *bytes, uint32_t const *byte_offsets, uint32_t n_strings,
                                   uint32_t *counts, uint32_t const *rune_offsets, uint32_t *runes) {
    uint32_t const index = blockIdx.x * blockDim.x + threadIdx.x;
    if (index >= n_strings) return;
    uint32_t const begin = byte_offsets[index], end = byte_offsets[index + 1];
    uint32_t produced = 0;
    uint32_t *const out = runes ? runes + rune_offsets[index] : nullptr;
    for (uint32_t i = begin; i < end;) {
        uint32_t const lead = bytes[i];
        uint32_t const length = 1u + (lead >= 0xC0u) + (lead >= 0xE0u) + (lead >= 0xF0u);
        if (out) {
            uint32_t const b1 = i + 1 < end ? bytes[i + 1] : 0, b2 = i + 2 < end ? bytes[i + 2] : 0,
                           b3 = i + 3 < end ? bytes[i + 3] : 0;
            uint32_t rune;
            if (length == 1) rune = lead;
            else if (length == 2) rune = (lead & 0x1Fu) << 6 | (b1 & 0x3Fu);
            else if (length == 3) rune = (lead & 0x0Fu) << 12 | (b1 & 0x3Fu) << 6 | (b2 & 0x3Fu);
            else rune = (lead & 0x07u) << 18 | (b1 & 0x3Fu) << 12 | (b2 & 0x3Fu) << 6 | (b3 & 0x3Fu);
            out[produced] = rune;
        }
        ++produced;
        i += length;
    }
    if (counts) counts[index] = produced;
}

/**
 *  Rune compaction for UTF-8 engines. When a call's strings draw on at most 256 distinct runes (any alphabetic script,
 *  small CJK corpora), ranking them into byte ids turns the rune problem into a byte problem and unlocks every
 *  byte-level kernel, including the bit-parallel ones. An open-addressing table in global memory collects the runes.
 */
constexpr size_t kRoleSwapRows = 32; ///< calls with fewer query rows than this (and more candidates) trade roles
constexpr uint32_t kRuneTableSlots = 1u << 16;
constexpr uint32_t kRuneEmpty = 0xFFFFFFFFu; // no rune: the unchecked decoder yields at most 21 bits

__device__ __forceinline__ uint32_t rune_slot(uint32_t rune) { return (rune * 2654435761u) >> 16; }

__global__ void rune_insert_kernel(uint32_t const *runes, uint64_t n, uint32_t *table, uint32_t *state) {
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t const rune = runes[i];
        uint32_t slot = rune_slot(rune);
        for (uint32_t probes = 0;; ++probes, slot = (slot + 1) & (kRuneTableSlots - 1)) {
            uint32_t const seen = table[slot];
            if (seen == rune) break;
            if (seen == kRuneEmpty) {
                uint32_t const previous = atomicCAS(&table[slot], kRuneEmpty, rune);
                if (previous == kRuneEmpty || previous == rune) break;
            }
            if (probes > 4096) { // hopelessly crowded: far more than 256 distinct runes anyway
                state[1] = 1;
                break;
            }
        }
    }
}

/** One block: ranks the occupied slots in slot order; state[0] = number of distinct runes. */
__global__ void rune_rank_kernel(uint32_t const *table, uint32_t *ranks, uint32_t *state) {
    __shared__ uint32_t partial[1024];
    constexpr uint32_t per_thread = kRuneTableSlots / 1024;
    uint32_t const first = threadIdx.x * per_thread;
    uint32_t mine = 0;
    for (uint32_t i = 0; i < per_thread; ++i) mine += table[first + i] != kRuneEmpty;
    partial[threadIdx.x] = mine;
    __syncthreads();
    for (uint32_t offset = 1; offset < 1024; offset <<= 1) { // Hillis-Steele inclusive scan
        uint32_t const add = threadIdx.x >= offset ? partial[threadIdx.x - offset] : 0;
        __syncthreads();
        partial[threadIdx.x] += add;
        __syncthreads();
    }
    uint32_t running = partial[threadIdx.x] - mine;
    for (uint32_t i = 0; i < per_thread; ++i)
        if (table[first + i] != kRuneEmpty) ranks[first + i] = running++;
    if (threadIdx.x == 1023) state[0] = partial[1023];
}

__global__ void rune_map_kernel(uint32_t const *runes, uint64_t n, uint32_t const *table, uint32_t const *ranks,
                                uint8_t *out) {
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t const rune = runes[i];
        uint32_t slot = rune_slot(rune);
        while (table[slot] != rune) slot = (slot + 1) & (kRuneTableSlots - 1);
        out[i] = (uint8_t)ranks[slot];
    }
}

#pragma endregion

#pragma region Errors

struct Failure {
    sz_status_t status = sz_success_k;
    char const *message = nullptr;
    explicit operator bool() const { return status != sz_success_k; }
};

static Failure ok() { return {}; }
static Failure fail(sz_status_t status, char const *message) { return {status, message}; }
static Failure fail_cuda(cudaError_t error) {
    // cudaGetErrorName returns a static string, like the reference's `cuGetErrorName` path (stringzillas.cuh:236-257).
    if (error == cudaErrorMemoryAllocation) return {sz_bad_alloc_k, cudaGetErrorName(error)};
    if (error == cudaErrorNoDevice || error == cudaErrorInsufficientDriver || error == cudaErrorInvalidDevice)
        return {sz_missing_gpu_k, cudaGetErrorName(error)};
    return {sz_status_unknown_k, cudaGetErrorName(error)};
}
#define SZB_CUDA(call)                                   \
    do {                                                 \
        cudaError_t error_ = (call);                     \
        if (error_ != cudaSuccess) {                     \
            cudaGetLastError();                          \
            return fail_cuda(error_);                    \
        }                                                \
    } while (0)

static sz_status_t report(Failure failure, char const **error_message) {
    if (error_message) *error_message = failure.status == sz_success_k ? nullptr : failure.message;
    return failure.status;
}

#pragma endregion

#pragma region Device memory

/** Grow-only device allocation, like the reference's engine-owned `cuda_cross_buffers` (cuda.cuh:1724-1775). */
struct DeviceBuffer {
    void *ptr = nullptr;
    size_t capacity = 0;
    DeviceBuffer() = default;
    DeviceBuffer(DeviceBuffer const &) = delete;
    DeviceBuffer &operator=(DeviceBuffer const &) = delete;
    ~DeviceBuffer() { release(); }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr, capacity = 0;
    }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= capacity) return cudaSuccess;
        if (ptr) cudaFree(ptr), ptr = nullptr, capacity = 0;
        size_t const rounded = ((bytes + bytes / 4 + 4095) / 4096) * 4096;
        cudaError_t error = cudaMalloc(&ptr, rounded);
        if (error == cudaSuccess) capacity = rounded;
        return error;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(ptr); }
};

/** Grow-only pinned host arena for the small per-call uploads (offsets, slots, jobs, query order). */
struct PinnedBuffer {
    void *ptr = nullptr;
    size_t capacity = 0;
    ~PinnedBuffer() { release(); }
    void release() {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr, capacity = 0;
    }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= capacity) return cudaSuccess;
        if (ptr) cudaFreeHost(ptr), ptr = nullptr, capacity = 0;
        size_t const rounded = ((bytes + bytes / 4 + 4095) / 4096) * 4096;
        cudaError_t error = cudaMallocHost(&ptr, rounded);
        if (error == cudaSuccess) capacity = rounded;
        return error;
    }
};

#pragma endregion

#pragma region Device scopes

/** Reference: `device_scope_t` = variant<default, cpu, gpu> (c/stringzillas/stringzillas.cuh:275-332). */
struct Scope {
    enum Kind { kDefault, kCpu, kGpu } kind = kDefault;
    size_t cpu_cores = 1;
    int gpu_device = 0;
    cudaStream_t stream = nullptr; // created lazily on first use, one stream per scope (types.cuh:403-473)
    std::mutex mutex;

    ~Scope() {
        if (stream) {
            cudaSetDevice(gpu_device);
            cudaStreamDestroy(stream);
        }
    }
    Failure ensure_stream() {
        std::lock_guard<std::mutex> lock(mutex);
        SZB_CUDA(cudaSetDevice(gpu_device));
        if (!stream) SZB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        return ok();
    }
};

static int visible_device_count() {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count;
}

/** Cached (sm_code, multiprocessor count) per device: `cudaGetDeviceProperties` costs milliseconds per call. */
struct DeviceFacts {
    int sm_code = 0, sms = 0;
};
static DeviceFacts device_facts(int device) {
    static std::mutex mutex;
    static DeviceFacts cache[64];
    static bool known[64] = {false};
    std::lock_guard<std::mutex> lock(mutex);
    if (device >= 0 && device < 64 && known[device]) return cache[device];
    DeviceFacts facts;
    int major = 0, minor = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) == cudaSuccess &&
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device) == cudaSuccess &&
        cudaDeviceGetAttribute(&facts.sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess)
        facts.sm_code = major * 10 + minor;
    else cudaGetLastError();
    if (device >= 0 && device < 64 && facts.sm_code) cache[device] = facts, known[device] = true;
    return facts;
}

/** Compute capability major*10+minor of a device, or 0. */
static int device_sm_code(int device) { return device_facts(device).sm_code; }

#pragma endregion

#pragma region Inputs

/** A string collection as the three ABI layouts present it, normalised to: bytes + 32-bit offsets. */
struct InputView {
    enum Layout { kSequence, kTape32, kTape64 } layout = kTape64;
    void const *object = nullptr;
    size_t count() const {
        switch (layout) {
        case kSequence: return (size_t) static_cast<sz_sequence_t const *>(object)->count;
        case kTape32: return (size_t) static_cast<sz_sequence_u32tape_t const *>(object)->count;
        default: return (size_t) static_cast<sz_sequence_u64tape_t const *>(object)->count;
        }
    }
};

enum class MemoryKind { kPageable, kPinned, kManaged, kDevice };

static MemoryKind memory_kind(void const *pointer) {
    if (!pointer) return MemoryKind::kPageable;
    cudaPointerAttributes attributes;
    if (cudaPointerGetAttributes(&attributes, pointer) != cudaSuccess) {
        cudaGetLastError();
        return MemoryKind::kPageable;
    }
    switch (attributes.type) {
    case cudaMemoryTypeHost: return MemoryKind::kPinned;
    case cudaMemoryTypeManaged: return MemoryKind::kManaged;
    case cudaMemoryTypeDevice: return MemoryKind::kDevice;
    default: return MemoryKind::kPageable;
    }
}

static bool pointer_is_host_readable(void const *pointer) { return memory_kind(pointer) != MemoryKind::kDevice; }

/** One side of the cross product: what the host knows about it, and where its bytes live on the device. */
struct StagedSet {
    size_t count = 0;
    uint64_t total_bytes = 0;
    std::vector<uint32_t> offsets; // count + 1, relative to the first byte of the collection
    uint32_t longest = 0;
    // Where the caller keeps the bytes.
    uint8_t const *blob = nullptr;           // tapes: first byte of the collection
    sz_sequence_t const *sequence = nullptr; // callback sequences
    bool host_readable = true;               // the bytes can be read by host code (pageable, pinned or managed memory)
    bool host_resident = true;               // ... and live in host memory proper (managed pages may sit on the device)
    bool pageable = false;                   // ... unregistered host memory: DMA needs a pinned bounce buffer
    uint8_t const *host_view = nullptr;      // the bytes as one host-readable run, when there is one
    bool uploaded = false;                   // the raw bytes are in `raw`
    DeviceBuffer raw, classes, device_offsets, gather_list;
    DeviceBuffer rune_counts, rune_offsets, runes; // UTF-8 engines: the decoded symbol streams
    std::vector<uint32_t> rune_counts_host;
    PinnedBuffer gather; // only for callback sequences
    std::vector<uint8_t> offsets_bounce;
    void release() {
        raw.release(), classes.release(), device_offsets.release(), gather_list.release(), rune_counts.release(), rune_offsets.release();
        runes.release(), gather.release();
    }
    /** Calls `visit(pointer, length, index)` for every string; only valid when `host_readable`. */
    template <typename visitor_type>
    void for_each_host_string(visitor_type &&visit) const {
        for (size_t i = 0; i < count; ++i) {
            size_t const length = offsets[i + 1] - offsets[i];
            uint8_t const *start = sequence ? (uint8_t const *)sequence->get_start(sequence->handle, i) : blob + offsets[i];
            visit(start, length, i);
        }
    }
};

#pragma endregion

#pragma region Engines

enum class Algorithm { kLevenshtein = 0, kNeedlemanWunsch = 1, kSmithWaterman = 2 };

struct Engine {
    Algorithm algorithm = Algorithm::kLevenshtein;
    // Costs exactly as the caller passed them (reference conventions).
    int match = 0, mismatch = 1, open = 1, extend = 1;
    uint8_t byte_to_class[256] = {0};
    int8_t costs[32 * 32] = {0};
    bool bit_parallel = true; // unit-cost Levenshtein may take the bit-parallel kernel (szs_b200_set_bit_parallel)
    bool utf8 = false;        // Levenshtein over UTF-8 runes instead of bytes (szs_levenshtein_distances_utf8*)

    // Grow-only device state, one set per engine: engines are not re-entrant.
    StagedSet queries, candidates;
    DeviceBuffer results, results_home, flags, query_flags, rune_table, rune_ranks, rune_state, scratch;
    // Everything the host generates for one call (offsets, scoring model, class map, query order, slots, jobs,
    // bit-parallel candidates, zeroed counters — and, for small host-resident inputs, the class-mapped strings
    // themselves) travels in ONE pinned arena and ONE host-to-device copy.
    PinnedBuffer arena_host, download;
    DeviceBuffer arena_device;
    cudaEvent_t events[4] = {nullptr, nullptr, nullptr, nullptr};
    // Independent launches of one call (different strip widths, word counts, families) overlap their tails on side
    // streams forked from, and joined back into, the scope's stream.
    static constexpr int kSideStreams = 3;
    cudaStream_t side_streams[kSideStreams] = {nullptr, nullptr, nullptr};
    cudaEvent_t fork_event = nullptr, join_events[kSideStreams] = {nullptr, nullptr, nullptr};
    int events_device = -1;
    szs_b200_call_stats_t stats = {};
    std::mutex mutex;
    // Multi-GPU calls (szs_b200_scores_u64tape_sharded) keep one clone of this engine per extra device.
    std::vector<std::unique_ptr<Engine>> shards;

    ~Engine() {
        if (events_device >= 0) cudaSetDevice(events_device);
        for (cudaEvent_t event : events)
            if (event) cudaEventDestroy(event);
        for (cudaEvent_t event : join_events)
            if (event) cudaEventDestroy(event);
        if (fork_event) cudaEventDestroy(fork_event);
        for (cudaStream_t stream : side_streams)
            if (stream) cudaStreamDestroy(stream);
    }
    /** A fresh engine with the same algorithm and costs (and no device state). */
    std::unique_ptr<Engine> clone() const {
        std::unique_ptr<Engine> copy(new (std::nothrow) Engine());
        if (!copy) return copy;
        copy->algorithm = algorithm, copy->match = match, copy->mismatch = mismatch, copy->open = open;
        copy->extend = extend, copy->bit_parallel = bit_parallel, copy->utf8 = utf8;
        std::memcpy(copy->byte_to_class, byte_to_class, sizeof(byte_to_class));
        std::memcpy(copy->costs, costs, sizeof(costs));
        return copy;
    }
};

#pragma endregion

#pragma region Staging

/**
 *  Reads the layout of one collection on the host: string count, 32-bit offsets relative to its first byte, the
 *  longest string, and where the bytes are. Copies nothing but (for device-resident tapes) the offsets.
 */
static Failure read_layout(InputView view, StagedSet &set, cudaStream_t stream) {
    size_t const count = view.count();
    set.count = count;
    set.offsets.resize(count + 1);
    set.longest = 0, set.blob = nullptr, set.sequence = nullptr, set.host_view = nullptr, set.uploaded = false;
    set.host_readable = true, set.host_resident = true, set.pageable = false;

    if (view.layout == InputView::kTape32 || view.layout == InputView::kTape64) {
        void const *offsets_pointer;
        size_t offset_width;
        uint8_t const *blob;
        if (view.layout == InputView::kTape32) {
            auto const *tape = static_cast<sz_sequence_u32tape_t const *>(view.object);
            blob = (uint8_t const *)tape->data, offsets_pointer = tape->offsets, offset_width = 4;
        }
        else {
            auto const *tape = static_cast<sz_sequence_u64tape_t const *>(view.object);
            blob = (uint8_t const *)tape->data, offsets_pointer = tape->offsets, offset_width = 8;
        }
        if (count == 0) { set.offsets[0] = 0, set.total_bytes = 0; }
        else {
            if (!offsets_pointer) return fail(sz_unexpected_dimensions_k, "Tape offsets are null");
            // Offsets are read on the host (as the reference does, cuda.cuh:4285-4295); bounce device-only ones.
            void const *host_offsets = offsets_pointer;
            if (!pointer_is_host_readable(offsets_pointer)) {
                set.offsets_bounce.resize((count + 1) * offset_width);
                SZB_CUDA(cudaMemcpyAsync(set.offsets_bounce.data(), offsets_pointer, (count + 1) * offset_width,
                                         cudaMemcpyDeviceToHost, stream));
                SZB_CUDA(cudaStreamSynchronize(stream));
                host_offsets = set.offsets_bounce.data();
            }
            // Every offset is validated as a 64-bit value before it is narrowed: monotone, and within 4 GB of the first.
            uint64_t first, previous;
            auto read = [&](size_t i) -> uint64_t {
                return offset_width == 4 ? (uint64_t) static_cast<uint32_t const *>(host_offsets)[i]
                                         : static_cast<uint64_t const *>(host_offsets)[i];
            };
            first = previous = read(0);
            set.offsets[0] = 0;
            for (size_t i = 1; i <= count; ++i) {
                uint64_t const current = read(i);
                if (current < previous) return fail(sz_unexpected_dimensions_k, "Tape offsets decrease");
                if (current - first > 0xFFFFFFF0ull) return fail(sz_overflow_risk_k, "Collection exceeds 4 GB of text");
                set.offsets[i] = (uint32_t)(current - first);
                set.longest = std::max(set.longest, (uint32_t)(current - previous));
                previous = current;
            }
            set.total_bytes = previous - first;
            blob += first;
        }
        set.blob = blob;
        MemoryKind const kind = set.total_bytes == 0 ? MemoryKind::kPageable : memory_kind(blob);
        set.host_readable = kind != MemoryKind::kDevice;
        set.host_resident = kind == MemoryKind::kPageable || kind == MemoryKind::kPinned;
        set.pageable = kind == MemoryKind::kPageable;
        if (set.host_readable) set.host_view = blob;
    }
    else {
        auto const *sequence = static_cast<sz_sequence_t const *>(view.object);
        set.sequence = sequence;
        uint64_t running = 0;
        for (size_t i = 0; i < count; ++i) {
            set.offsets[i] = (uint32_t)running;
            uint64_t const length = sequence->get_length(sequence->handle, i);
            running += length;
            if (running > 0xFFFFFFF0ull) return fail(sz_overflow_risk_k, "Collection exceeds 4 GB of text");
            set.longest = std::max(set.longest, (uint32_t)length);
        }
        set.offsets[count] = (uint32_t)running;
        set.total_bytes = running;
        // Like the reference (cuda.cuh:4268-4272), only the first string's pointer is inspected.
        MemoryKind const kind =
            (count == 0 || running == 0) ? MemoryKind::kPageable : memory_kind(sequence->get_start(sequence->handle, 0));
        set.host_readable = kind != MemoryKind::kDevice;
        set.host_resident = kind == MemoryKind::kPageable || kind == MemoryKind::kPinned;
    }
    return ok();
}

/** Pageable blobs of at least this many bytes are staged through pinned memory by several host threads. */
constexpr uint64_t kParallelStagingBytes = 4ull << 20;

/** Copies the raw bytes of one collection to the device (`set.raw`). Accepts host, pinned, managed or device memory. */
static Failure upload_raw(StagedSet &set, cudaStream_t stream, szs_b200_call_stats_t &stats) {
    // Padding keeps the 16-byte vector loads of the class-mapping kernel in bounds.
    SZB_CUDA(set.raw.reserve(set.total_bytes + 256));
    SZB_CUDA(set.classes.reserve(set.total_bytes + 256));
    set.uploaded = true;
    if (!set.total_bytes) return ok();
    if (!set.sequence) {
        if (set.pageable && set.total_bytes >= kParallelStagingBytes) {
            // A large blob in pageable memory (a retrieval corpus): the driver would bounce it through its own pinned
            // buffer with one host thread. A few host threads copy it into the engine's pinned buffer chunk by chunk while
            // the copy engine drains the finished chunks.
            SZB_CUDA(set.gather.reserve(set.total_bytes));
            uint64_t const chunk = 1ull << 20, n_chunks = (set.total_bytes + chunk - 1) / chunk;
            std::vector<std::atomic<uint8_t>> done(n_chunks);
            for (auto &flag : done) flag.store(0, std::memory_order_relaxed);
            std::atomic<uint64_t> next {0};
            unsigned const n_threads = (unsigned)std::min<uint64_t>(std::max(2u, std::min(6u, std::thread::hardware_concurrency() / 2)), n_chunks);
            auto worker = [&] {
                for (uint64_t i = next.fetch_add(1); i < n_chunks; i = next.fetch_add(1)) {
                    uint64_t const begin = i * chunk, bytes = std::min(chunk, set.total_bytes - begin);
                    std::memcpy((char *)set.gather.ptr + begin, set.blob + begin, bytes);
                    done[i].store(1, std::memory_order_release);
                }
            };
            std::vector<std::thread> threads;
            try {
                for (unsigned t = 0; t + 1 < n_threads; ++t) threads.emplace_back(worker);
            }
            catch (...) { // no helper threads to be had: the loop below copies every chunk itself
            }
            cudaError_t error = cudaSuccess;
            for (uint64_t i = 0; i < n_chunks; ++i) {
                while (!done[i].load(std::memory_order_acquire)) {
                    // Help with the copying instead of spinning.
                    uint64_t const mine = next.fetch_add(1);
                    if (mine < n_chunks) {
                        uint64_t const begin = mine * chunk, bytes = std::min(chunk, set.total_bytes - begin);
                        std::memcpy((char *)set.gather.ptr + begin, set.blob + begin, bytes);
                        done[mine].store(1, std::memory_order_release);
                    }
                    else std::this_thread::yield();
                }
                uint64_t const begin = i * chunk, bytes = std::min(chunk, set.total_bytes - begin);
                if (error == cudaSuccess)
                    error = cudaMemcpyAsync((char *)set.raw.ptr + begin, (char *)set.gather.ptr + begin, bytes, cudaMemcpyHostToDevice, stream);
            }
            for (std::thread &thread : threads) thread.join();
            SZB_CUDA(error);
            stats.h2d_bytes += set.total_bytes;
            set.host_view = (uint8_t const *)set.gather.ptr;
            return ok();
        }
        SZB_CUDA(cudaMemcpyAsync(set.raw.ptr, set.blob, set.total_bytes, cudaMemcpyDefault, stream));
        if (set.host_readable) stats.h2d_bytes += set.total_bytes;
        return ok();
    }
    sz_sequence_t const *sequence = set.sequence;
    if (set.host_resident) {
        SZB_CUDA(set.gather.reserve(set.total_bytes));
        set.for_each_host_string([&](uint8_t const *start, size_t length, size_t i) {
            if (length) std::memcpy((char *)set.gather.ptr + set.offsets[i], start, length);
        });
        SZB_CUDA(cudaMemcpyAsync(set.raw.ptr, set.gather.ptr, set.total_bytes, cudaMemcpyHostToDevice, stream));
        stats.h2d_bytes += set.total_bytes;
        set.host_view = (uint8_t const *)set.gather.ptr;
        return ok();
    }
    // Managed or device memory: the strings are gathered ON the device from an uploaded pointer list (the host never
    // touches managed pages that may sit on the device, and device strings need no per-string copy calls).
    size_t const pointer_bytes = set.count * sizeof(void *), offset_bytes = (set.count + 1) * sizeof(uint32_t);
    SZB_CUDA(set.gather.reserve(pointer_bytes + offset_bytes));
    SZB_CUDA(set.gather_list.reserve(pointer_bytes + offset_bytes));
    void const **const pointers = (void const **)set.gather.ptr;
    for (size_t i = 0; i < set.count; ++i) pointers[i] = sequence->get_start(sequence->handle, i);
    std::memcpy((char *)set.gather.ptr + pointer_bytes, set.offsets.data(), offset_bytes);
    SZB_CUDA(cudaMemcpyAsync(set.gather_list.ptr, set.gather.ptr, pointer_bytes + offset_bytes, cudaMemcpyHostToDevice, stream));
    stats.h2d_bytes += pointer_bytes + offset_bytes;
    int const blocks = (int)std::min<size_t>((set.count + 7) / 8, 4096);
    SZB_LAUNCH(gather_strings_kernel, blocks, 256, 0, stream, (uint8_t const *const *)set.gather_list.ptr,
               (uint32_t const *)((char *)set.gather_list.ptr + pointer_bytes), (uint32_t)set.count,
               set.raw.as<uint8_t>());
    SZB_CUDA(cudaGetLastError());
    stats.kernel_launches++;
    return ok();
}

/** The per-call arena: sections are carved first (sizes only), then both copies are reserved once, then filled. */
struct ArenaLayout {
    size_t used = 0;
    size_t carve(size_t bytes) {
        size_t const at = used;
        used = (used + bytes + 255) & ~(size_t)255;
        return at;
    }
};

#pragma endregion

#pragma region Scheduling

/** Everything the host decides about one call's scoring model. */
struct Model {
    Scoring scoring;        // uploaded verbatim
    uint8_t lut[256];       // byte -> class (identity in compare mode)
    int family16 = -1;      // packed family, or -1 when the cost signs rule the packed recurrences out
    int family32 = 0;       // 32-bit family, always available
    int64_t range_offset = 0;   // 16-bit tier: constant part of the largest cell value
    int64_t range_per_step = 0; // 16-bit tier: growth per min(query, candidate) character
    bool symmetric_costs = true;
};

/**
 *  Translates the engine's costs into the kernels' maximising convention and decides which recurrence families
 *  are exact for them. `present` flags the byte values occurring among the candidates (Levenshtein only).
 */
static void build_model(Engine const &engine, uint32_t const *present, bool wide_symbols, Model &model) {
    Scoring &sc = model.scoring;
    std::memset(&sc, 0, sizeof(sc));
    bool const local = engine.algorithm == Algorithm::kSmithWaterman;
    bool linear = engine.open == engine.extend;
    int smin = 127, smax = -128;

    if (engine.algorithm == Algorithm::kLevenshtein) {
        // Distances are minimised: negate every cost, maximise, negate the result.
        sc.negate = 1;
        sc.open = -engine.open, sc.extend = -engine.extend;
        sc.match = -engine.match, sc.mismatch = -engine.mismatch;
        int distinct = 0;
        for (int b = 0; b < 256; ++b) distinct += present[b] != 0;
        if (distinct <= 31 && !wide_symbols) {
            int next = 0;
            for (int b = 0; b < 256; ++b) model.lut[b] = present[b] ? (uint8_t)next++ : (uint8_t)distinct;
            sc.n_classes = distinct + 1; // the last class collects query bytes no candidate contains
            for (int a = 0; a < 32; ++a)
                for (int b = 0; b < 32; ++b)
                    sc.sub[a * 32 + b] = (int8_t)((a == b && a < distinct) ? sc.match : sc.mismatch);
            smin = std::min(sc.match, sc.mismatch), smax = std::max(sc.match, sc.mismatch);
        }
        else {
            for (int b = 0; b < 256; ++b) model.lut[b] = (uint8_t)b;
            sc.compare = 1, sc.n_classes = 0;
            smin = std::min(sc.match, sc.mismatch), smax = std::max(sc.match, sc.mismatch);
        }
    }
    else {
        sc.open = engine.open, sc.extend = engine.extend;
        int top = 0;
        for (int b = 0; b < 256; ++b) model.lut[b] = engine.byte_to_class[b] & 31, top = std::max<int>(top, model.lut[b]);
        sc.n_classes = top + 1;
        std::memcpy(sc.sub, engine.costs, sizeof(sc.sub));
        for (int a = 0; a <= top; ++a)
            for (int b = 0; b <= top; ++b) {
                smin = std::min<int>(smin, sc.sub[a * 32 + b]), smax = std::max<int>(smax, sc.sub[a * 32 + b]);
                if (sc.sub[a * 32 + b] != sc.sub[b * 32 + a]) model.symmetric_costs = false;
            }
    }

    model.family32 = local ? (linear ? FAM_S32_LOCAL_LIN : FAM_S32_LOCAL_AFF)
                           : (linear ? FAM_S32_GLOBAL_LIN : FAM_S32_GLOBAL_AFF);
    model.family16 = -1;
    if (sc.compare) {
        // Large raw-byte alphabets: the packed tier has a byte-compare flavour of the shifted linear recurrence only;
        // it needs a match to score at least as well as a mismatch (the XOR trick subtracts a non-negative drop).
        if (!wide_symbols && linear && sc.match >= sc.mismatch) {
            model.family16 = FAM_LIN_COMPARE;
            model.range_offset = 0, model.range_per_step = std::max(smax - 2 * sc.open, 0);
        }
        return;
    }

    int const open = sc.open, extend = sc.extend;
    if (!local && linear) {
        // u = H - (i + j) * g, e = max(S - 2g, 0): exact for every sign of g. Cells stay below min(m, n) * e_max.
        model.family16 = FAM_LIN_SHIFT;
        sc.bias = 0;
        model.range_offset = 0, model.range_per_step = std::max(smax - 2 * open, 0);
    }
    else if (!local && extend <= 0 && extend - open >= 0) {
        // The finite "discard" seeds of the reference's Gotoh baselines equal -inf seeds iff extend <= 0.
        int const go = extend - open;
        model.family16 = FAM_AFF_SHIFT;
        sc.bias = 4 * go;
        model.range_offset = 4 * go, model.range_per_step = std::max(std::max(smax, 2 * open) - 2 * extend, 0);
    }
    else if (local && linear && open <= 0) {
        model.family16 = FAM_SW_LIN;
        sc.bias = std::max(std::max(-smin, -open), 0);
        model.range_offset = sc.bias, model.range_per_step = std::max(smax, 0);
    }
    else if (local && !linear && open <= 0 && extend <= 0) {
        model.family16 = FAM_SW_AFF;
        sc.bias = std::max(std::max(-smin, -(open + 2 * extend)), std::max(-open, 0));
        model.range_offset = sc.bias, model.range_per_step = std::max(smax, 0);
    }
}

/** Whether every cell (and the retired intermediate of one more step) of a pair fits an unsigned 16-bit half. */
static bool fits_u16(Model const &model, uint64_t shorter) {
    if (model.family16 < 0) return false;
    return (uint64_t)model.range_offset + (shorter + 2) * (uint64_t)model.range_per_step <= 65000ull;
}

/** Whether the 32-bit tier is exact: |score| <= (m + n + 4) * magnitude must stay far from 2^31 (and the 2^28 pad). */
static bool fits_s32(Engine const &engine, uint64_t total_length) {
    int magnitude = 1;
    if (engine.algorithm == Algorithm::kLevenshtein)
        magnitude = std::max(std::max(std::abs(engine.match), std::abs(engine.mismatch)),
                             std::max(std::abs(engine.open), std::abs(engine.extend)));
    else {
        magnitude = std::max(std::abs(engine.open), std::abs(engine.extend));
        for (int i = 0; i < 32 * 32; ++i) magnitude = std::max<int>(magnitude, std::abs(engine.costs[i]));
    }
    magnitude = std::max(magnitude, 1);
    return (total_length + 4) * (uint64_t)magnitude < (1ull << 27);
}

struct LaunchPlan {
    int family, strip_index;
    uint32_t first_job, n_jobs;
};

#pragma endregion

#pragma region The cross product

static Failure ensure_events(Engine &engine, int device) {
    if (engine.events_device == device && engine.events[0]) return ok();
    if (engine.events_device >= 0 && engine.events_device != device) {
        // The engine moves to another GPU (a scope of a different device): its grow-only buffers live on the old one.
        cudaSetDevice(engine.events_device);
        engine.queries.release(), engine.candidates.release();
        for (DeviceBuffer *buffer : {&engine.results, &engine.results_home, &engine.flags, &engine.query_flags,
                                     &engine.rune_table, &engine.rune_ranks, &engine.rune_state, &engine.scratch,
                                     &engine.arena_device})
            buffer->release();
        engine.arena_host.release(), engine.download.release();
        cudaSetDevice(device);
    }
    for (cudaEvent_t &event : engine.events) {
        if (event) cudaEventDestroy(event), event = nullptr;
        SZB_CUDA(cudaEventCreate(&event));
    }
    for (int i = 0; i < Engine::kSideStreams; ++i) {
        if (engine.side_streams[i]) cudaStreamDestroy(engine.side_streams[i]), engine.side_streams[i] = nullptr;
        if (engine.join_events[i]) cudaEventDestroy(engine.join_events[i]), engine.join_events[i] = nullptr;
        SZB_CUDA(cudaStreamCreateWithFlags(&engine.side_streams[i], cudaStreamNonBlocking));
        SZB_CUDA(cudaEventCreateWithFlags(&engine.join_events[i], cudaEventDisableTiming));
    }
    if (engine.fork_event) cudaEventDestroy(engine.fork_event), engine.fork_event = nullptr;
    SZB_CUDA(cudaEventCreateWithFlags(&engine.fork_event, cudaEventDisableTiming));
    engine.events_device = device;
    return ok();
}

/** Strided 2-D copy of `rows` x `width` bytes that degenerates to ONE linear copy when both sides are dense: the copy
 *  engine pays per row, and a score matrix has thousands of short rows. */
static cudaError_t copy_rows(void *destination, size_t destination_pitch, void const *source, size_t source_pitch,
                             size_t width, size_t rows, cudaMemcpyKind kind, cudaStream_t stream) {
    if (destination_pitch == width && source_pitch == width)
        return cudaMemcpyAsync(destination, source, width * rows, kind, stream);
    return cudaMemcpy2DAsync(destination, destination_pitch, source, source_pitch, width, rows, kind, stream);
}

/** Grid of a grid-stride helper kernel: enough blocks for `items`, capped at a few waves of the device. */
static int helper_grid(uint64_t items, int per_block, int device_sms, int waves) {
    return (int)std::max<uint64_t>(1, std::min<uint64_t>((items + per_block - 1) / per_block, (uint64_t)device_sms * waves));
}

/**
 *  @brief Steps 2-2b of a call: which symbols are we scoring?
 *
 *  Levenshtein engines learn which byte values the candidates contain (`present`), so that the alphabet can be compacted
 *  into classes. UTF-8 engines additionally decide between three representations: pure ASCII stays bytes; at most 256
 *  distinct runes are ranked into byte ids (the staged byte streams and offsets are REPLACED by the id streams);
 *  anything else becomes UTF-32 symbol streams (`wide_symbols`), with offsets and lengths counted in symbols.
 *
 *  `packed` (small host-resident inputs: nothing has been uploaded yet, the host scans the caller's memory) is given up
 *  here when a UTF-8 engine meets non-ASCII input: the rune machinery lives on the device.
 */
static Failure resolve_symbols(Engine &engine, cudaStream_t stream, bool symmetric, bool &packed, int device_sms,
                               StagedSet &qs, StagedSet &cs, szs_b200_call_stats_t &stats, uint32_t (&present)[256],
                               bool &wide_symbols) {
    // 2. The scoring model; Levenshtein compacts the alphabet of the candidates into <= 31 classes.
    if (engine.algorithm == Algorithm::kLevenshtein) {
        uint32_t query_present[256] = {0};
        bool const need_queries = engine.utf8 && !symmetric && qs.total_bytes;
        // Small host-resident inputs are scanned on the host: a kernel plus a D2H round-trip costs more than the scan.
        constexpr uint64_t host_scan_limit = 512 * 1024;
        auto scannable = [&](StagedSet const &set) {
            return !set.total_bytes || (set.total_bytes <= host_scan_limit && set.host_resident && (set.host_view || packed));
        };
        if (scannable(cs) && (!need_queries || scannable(qs))) {
            auto scan = [](StagedSet const &set, uint32_t (&seen)[256]) {
                if (set.host_view) mark_present_bytes(set.host_view, set.total_bytes, seen);
                else
                    set.for_each_host_string([&](uint8_t const *start, size_t length, size_t) { mark_present_bytes(start, length, seen); });
            };
            scan(cs, present);
            if (need_queries) scan(qs, query_present);
        }
        else {
            SZB_CUDA(engine.flags.reserve(256 * sizeof(uint32_t)));
            SZB_CUDA(cudaMemsetAsync(engine.flags.ptr, 0, 256 * sizeof(uint32_t), stream));
            if (cs.total_bytes) {
                SZB_LAUNCH(byte_presence_kernel, 2 * device_sms, 512, 0, stream, cs.raw.as<uint8_t>(), cs.total_bytes,
                           engine.flags.as<uint32_t>());
                SZB_CUDA(cudaGetLastError());
                stats.kernel_launches++;
            }
            if (need_queries) {
                SZB_CUDA(engine.query_flags.reserve(256 * sizeof(uint32_t)));
                SZB_CUDA(cudaMemsetAsync(engine.query_flags.ptr, 0, 256 * sizeof(uint32_t), stream));
                SZB_LAUNCH(byte_presence_kernel, 2 * device_sms, 512, 0, stream, qs.raw.as<uint8_t>(), qs.total_bytes,
                           engine.query_flags.as<uint32_t>());
                SZB_CUDA(cudaGetLastError());
                stats.kernel_launches++;
                SZB_CUDA(cudaMemcpyAsync(query_present, engine.query_flags.ptr, sizeof(query_present),
                                         cudaMemcpyDeviceToHost, stream));
                stats.d2h_bytes += sizeof(query_present);
            }
            SZB_CUDA(cudaMemcpyAsync(present, engine.flags.ptr, sizeof(present), cudaMemcpyDeviceToHost, stream));
            SZB_CUDA(cudaStreamSynchronize(stream));
            stats.d2h_bytes += sizeof(present);
        }
        if (engine.utf8)
            for (int b = 128; b < 256; ++b) wide_symbols |= present[b] != 0 || query_present[b] != 0;
    }
    if (!wide_symbols) return ok();

    // 2a. UTF-8 engines: pure-ASCII input is byte-level Levenshtein (as the reference's per-pair ASCII fallback,
    //     serial.hpp:2811-2814); anything else is decoded into UTF-32 symbol streams that replace the byte streams
    //     in everything downstream (lengths, offsets, scheduling), and scored by the symbol-compare recurrences.
    if (packed) {
        packed = false;
        if (Failure failure = upload_raw(qs, stream, stats)) return failure;
        if (!symmetric)
            if (Failure failure = upload_raw(cs, stream, stats)) return failure;
    }
    auto decode_set = [&](StagedSet &set) -> Failure {
        uint32_t const count = (uint32_t)set.count;
        SZB_CUDA(set.device_offsets.reserve((size_t)(count + 1) * sizeof(uint32_t)));
        SZB_CUDA(set.rune_counts.reserve((size_t)(count + 1) * sizeof(uint32_t)));
        SZB_CUDA(set.rune_offsets.reserve((size_t)(count + 1) * sizeof(uint32_t)));
        SZB_CUDA(cudaMemcpyAsync(set.device_offsets.ptr, set.offsets.data(), (size_t)(count + 1) * sizeof(uint32_t),
                                 cudaMemcpyHostToDevice, stream));
        stats.h2d_bytes += (size_t)(count + 1) * sizeof(uint32_t);
        int const blocks = (int)((count + 127) / 128);
        if (count) {
            SZB_LAUNCH(utf8_decode_kernel, blocks, 128, 0, stream, set.raw.as<uint8_t>(),
                       set.device_offsets.as<uint32_t>(), count, set.rune_counts.as<uint32_t>(), nullptr, nullptr);
            SZB_CUDA(cudaGetLastError());
            stats.kernel_launches++;
        }
        set.rune_counts_host.resize(count + 1);
        SZB_CUDA(cudaMemcpyAsync(set.rune_counts_host.data(), set.rune_counts.ptr, (size_t)count * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, stream));
        SZB_CUDA(cudaStreamSynchronize(stream)); // also: `set.offsets` is free to change from here on
        stats.d2h_bytes += (size_t)count * sizeof(uint32_t);
        uint64_t running = 0;
        set.longest = 0;
        for (uint32_t i = 0; i < count; ++i) {
            uint32_t const runes = set.rune_counts_host[i];
            set.offsets[i] = (uint32_t)running; // from here on, offsets and lengths are in symbols
            running += runes;
            set.longest = std::max(set.longest, runes);
        }
        set.offsets[count] = (uint32_t)running;
        SZB_CUDA(cudaMemcpyAsync(set.rune_offsets.ptr, set.offsets.data(), (size_t)(count + 1) * sizeof(uint32_t),
                                 cudaMemcpyHostToDevice, stream));
        SZB_CUDA(set.runes.reserve((running + 64) * sizeof(uint32_t)));
        if (count) {
            SZB_LAUNCH(utf8_decode_kernel, blocks, 128, 0, stream, set.raw.as<uint8_t>(),
                       set.device_offsets.as<uint32_t>(), count, nullptr, set.rune_offsets.as<uint32_t>(),
                       set.runes.as<uint32_t>());
            SZB_CUDA(cudaGetLastError());
            stats.kernel_launches++;
        }
        SZB_CUDA(cudaStreamSynchronize(stream)); // `set.offsets` must not change while the copy is in flight
        return ok();
    };
    if (Failure failure = decode_set(qs)) return failure;
    if (!symmetric)
        if (Failure failure = decode_set(cs)) return failure;

    // 2b. At most 256 distinct runes in the whole call? Then rank them into byte ids: the rune problem becomes a
    //     byte problem over a private alphabet, and every byte-level kernel applies (bit-parallel ones included).
    SZB_CUDA(engine.rune_table.reserve(kRuneTableSlots * sizeof(uint32_t)));
    SZB_CUDA(engine.rune_ranks.reserve(kRuneTableSlots * sizeof(uint32_t)));
    SZB_CUDA(engine.rune_state.reserve(2 * sizeof(uint32_t)));
    SZB_CUDA(cudaMemsetAsync(engine.rune_table.ptr, 0xFF, kRuneTableSlots * sizeof(uint32_t), stream));
    SZB_CUDA(cudaMemsetAsync(engine.rune_state.ptr, 0, 2 * sizeof(uint32_t), stream));
    auto for_each_side = [&](auto &&apply) -> Failure {
        if (Failure failure = apply(qs)) return failure;
        if (!symmetric)
            if (Failure failure = apply(cs)) return failure;
        return ok();
    };
    if (Failure failure = for_each_side([&](StagedSet &set) -> Failure {
            uint64_t const n = set.offsets[set.count];
            if (!n) return ok();
            SZB_LAUNCH(rune_insert_kernel, helper_grid(n, 256, device_sms, 8), 256, 0, stream,
                       set.runes.as<uint32_t>(), n, engine.rune_table.as<uint32_t>(),
                       engine.rune_state.as<uint32_t>());
            SZB_CUDA(cudaGetLastError());
            stats.kernel_launches++;
            return ok();
        }))
        return failure;
    SZB_LAUNCH(rune_rank_kernel, 1, 1024, 0, stream, engine.rune_table.as<uint32_t>(),
               engine.rune_ranks.as<uint32_t>(), engine.rune_state.as<uint32_t>());
    SZB_CUDA(cudaGetLastError());
    stats.kernel_launches++;
    uint32_t rune_state[2] = {0, 0};
    SZB_CUDA(cudaMemcpyAsync(rune_state, engine.rune_state.ptr, sizeof(rune_state), cudaMemcpyDeviceToHost, stream));
    SZB_CUDA(cudaStreamSynchronize(stream));
    stats.d2h_bytes += sizeof(rune_state);
    if (rune_state[1] == 0 && rune_state[0] <= 256) {
        if (Failure failure = for_each_side([&](StagedSet &set) -> Failure {
                uint64_t const n = set.offsets[set.count];
                // The id stream overwrites the staged UTF-8 bytes (never longer than them); the symbol offsets (already
                // in `set.offsets`) become the byte offsets of the byte-level problem.
                if (n) {
                    SZB_LAUNCH(rune_map_kernel, helper_grid(n, 256, device_sms, 8), 256, 0, stream,
                               set.runes.as<uint32_t>(), n, engine.rune_table.as<uint32_t>(),
                               engine.rune_ranks.as<uint32_t>(), set.raw.as<uint8_t>());
                    SZB_CUDA(cudaGetLastError());
                    stats.kernel_launches++;
                }
                set.total_bytes = n;
                set.host_view = nullptr;
                return ok();
            }))
            return failure;
        wide_symbols = false;
        std::memset(present, 0, sizeof(present));
        SZB_CUDA(engine.flags.reserve(256 * sizeof(uint32_t)));
        SZB_CUDA(cudaMemsetAsync(engine.flags.ptr, 0, 256 * sizeof(uint32_t), stream));
        if (cs.total_bytes) {
            SZB_LAUNCH(byte_presence_kernel, 2 * device_sms, 512, 0, stream, cs.raw.as<uint8_t>(), cs.total_bytes,
                       engine.flags.as<uint32_t>());
            SZB_CUDA(cudaGetLastError());
            stats.kernel_launches++;
        }
        SZB_CUDA(cudaMemcpyAsync(present, engine.flags.ptr, sizeof(present), cudaMemcpyDeviceToHost, stream));
        SZB_CUDA(cudaStreamSynchronize(stream));
        stats.d2h_bytes += sizeof(present);
    }
    return ok();
}

/** `SZB200_TRACE=1` prints host-side wall-clock marks of one call to stderr (diagnostics only). */
struct Trace {
    bool enabled = std::getenv("SZB200_TRACE") != nullptr;
    std::chrono::steady_clock::time_point start = std::chrono::steady_clock::now();
    void mark(char const *what) const {
        if (!enabled) return;
        double const us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - start).count();
        std::fprintf(stderr, "[szb200] %8.1f us  %s\n", us, what);
    }
};

/** A candidate pair (or single) of the DP tiers with the recurrence family and strip shape chosen for it. */
struct HostSlot {
    CandSlot slot;
    int family;
    Shape shape;
};

/** The candidates one bit-parallel launch covers (all with the same number of pattern words per lane). */
struct MyersLaunch {
    int words;
    std::vector<MyersCandidate> members; // longest first
};

/** Everything the host decides about one call before anything is launched; pure host data (unit-testable). */
struct Schedule {
    std::vector<uint32_t> order;    ///< query rows of this call, longest first
    std::vector<uint32_t> position; ///< inverse of `order` (symmetric triangle only)
    bool triangle = false;          ///< symmetric call under symmetric costs: each unordered pair scored once
    std::vector<HostSlot> slots;    ///< DP tiers
    std::vector<Job> jobs;          ///< DP tiers: slot x block of the query order, grouped by plan
    std::vector<LaunchPlan> plans;  ///< DP tiers: one launch per (family, strip width)
    uint64_t scratch_words_per_cta = 0;
    std::vector<MyersLaunch> myers_launches; ///< bit-parallel tier, heaviest launch first
    bool myers_shared_table = false; ///< one match table per CTA (raw bytes or runes) instead of one per warp
    bool myers_rune_table = false;   ///< ... whose rows are keyed by UTF-32 rune through a hash table
    int myers_rows = 0;
    uint64_t cells_total = 0, cells_u16 = 0, cells_myers = 0;
};

/** Bit-parallel launches of one call (different words per lane) share the SMs' resident CTA slots in proportion to
 *  their work and run side by side from the start, instead of each asking for every slot and queueing behind the other:
 *  a sweep is a coarse unit (c2: 2.7 per warp over both launches), so two launches in turn mean two ragged ends at a
 *  third of the occupancy, side by side they mean one. */
#ifndef SZB_SHARE_RESIDENT_SLOTS
#define SZB_SHARE_RESIDENT_SLOTS 0
#endif

/** EXPERIMENT, off by default: bit-parallel shapes are widened (more lanes, fewer words per lane: finer sweeps) until a
 *  call holds at least this many sweeps per resident warp. */
#ifndef SZB_MYERS_MIN_SWEEPS_PER_WARP
#define SZB_MYERS_MIN_SWEEPS_PER_WARP 0
#endif

/** Multi-pass slots park one boundary column per query row in per-CTA scratch: a job's block is cut so that its
 *  scratch stays below this many 32-bit words (16 MB), whatever the row count of the call. */
constexpr uint64_t kScratchWordsPerJob = 4ull << 20;

/**
 *  @brief Steps 4-6 of a call: order the queries, split the candidates between the bit-parallel tier and the DP tiers,
 *         pair DP candidates into slots, choose shapes, and cut everything into jobs / tasks. Host-only.
 *  @param n_columns candidates [0, n_columns) take part (all of them, except in a row shard of a symmetric call).
 */
static Failure build_schedule(Engine const &engine, Model const &model, StagedSet const &qs, StagedSet const &cs,
                              bool symmetric, bool whole_matrix, size_t row_begin, size_t n_rows, size_t n_columns,
                              uint32_t const (&present)[256], bool wide_symbols, int device_sms, Trace const &trace,
                              Schedule &schedule) {
    size_t const n_queries = qs.count, n_candidates = n_columns;
    // 4. Query order: longest first, so that the batches a warp sweeps together have similar lengths and the heavy
    //    jobs start early.
    auto query_length = [&](uint32_t q) { return qs.offsets[q + 1] - qs.offsets[q]; };
    std::vector<uint32_t> &order = schedule.order;
    order.assign(n_rows, 0);
    {   // One 64-bit key per string — (inverted length, index) — sorts stably with a plain integer sort.
        std::vector<uint64_t> keys(n_rows);
        for (size_t i = 0; i < n_rows; ++i) {
            uint32_t const q = (uint32_t)(row_begin + i);
            keys[i] = ((uint64_t)(0xFFFFFFFFu - query_length(q)) << 32) | q;
        }
        std::sort(keys.begin(), keys.end());
        for (size_t i = 0; i < n_rows; ++i) order[i] = (uint32_t)keys[i];
    }
    uint32_t const longest_query = query_length(order[0]);

    // Symmetric self-similarity under a symmetric cost model: score every unordered pair once, with the string that
    // comes first in the longest-first order as the query. A candidate at position p then needs only the queries at
    // positions [0, p] — a prefix of the order — and a ranked mirror fills the other half. (An asymmetric substitution
    // table must follow the reference's (query = i, candidate = j <= i) orientation: full matrix + index mirror.)
    bool const triangle = schedule.triangle = symmetric && whole_matrix && model.symmetric_costs && n_queries > 1;
    std::vector<uint32_t> &position = schedule.position;
    position.clear();
    if (triangle) {
        position.resize(n_queries);
        for (uint32_t p = 0; p < (uint32_t)n_rows; ++p) position[order[p]] = p;
    }
    auto rows_needed = [&](uint32_t candidate) { return triangle ? position[candidate] + 1 : (uint32_t)n_rows; };

    trace.mark("queries ordered");
    // 5. Slots: candidates sorted by length; neighbours share a register in the packed tier.
    auto candidate_length = [&](uint32_t c) { return cs.offsets[c + 1] - cs.offsets[c]; };
    std::vector<uint32_t> by_length(n_candidates);
    {
        std::vector<uint64_t> keys(n_candidates);
        for (size_t c = 0; c < n_candidates; ++c) keys[c] = ((uint64_t)candidate_length((uint32_t)c) << 32) | (uint32_t)c;
        std::sort(keys.begin(), keys.end());
        for (size_t c = 0; c < n_candidates; ++c) by_length[c] = (uint32_t)keys[c];
    }
    uint64_t query_bytes_in_rows = 0;
    for (uint32_t q : order) query_bytes_in_rows += query_length(q);

    // A call too small to give every warp of the device a pair of its own is bound by the latency of one sweep, not by
    // throughput: spread each pair over as many lanes as there are idle ones (more lanes = fewer columns per lane =
    // a shorter dependent chain per step).
    uint64_t const device_warps = (uint64_t)device_sms * (kThreads / 32);
    uint64_t const n_pairs = (uint64_t)n_rows * n_candidates;
    uint32_t latency_lanes = 1;
    while (latency_lanes < 32 && n_pairs * latency_lanes * 2 <= device_warps * 32) latency_lanes *= 2;
    uint32_t myers_min_lanes = latency_lanes;
    if (SZB_MYERS_MIN_SWEEPS_PER_WARP)
        while (myers_min_lanes < 32 && n_pairs * myers_min_lanes < (uint64_t)SZB_MYERS_MIN_SWEEPS_PER_WARP * 2 * device_warps * 32)
            myers_min_lanes *= 2;

    // 5a. Unit-cost Levenshtein over a compact alphabet: candidates the bit-parallel kernel covers leave the DP tiers.
    std::vector<MyersLaunch> &myers_launches = schedule.myers_launches;
    myers_launches.clear();
    uint64_t &cells_myers = schedule.cells_myers = 0;
    bool const unit_costs = engine.algorithm == Algorithm::kLevenshtein && engine.match == 0 && engine.mismatch == 1 &&
                            engine.open == 1 && engine.extend == 1;
    // Compact alphabets keep one match table per WARP (rows = classes); raw-byte alphabets of more than 31 symbols keep
    // one table per CTA whose rows are the bytes the candidate contains (bounded by the bytes any candidate contains).
    bool const myers_shared_table = schedule.myers_shared_table = model.scoring.compare != 0;
    int &myers_rows = schedule.myers_rows = model.scoring.n_classes;
    bool const myers_rune_table = schedule.myers_rune_table = wide_symbols;
    if (myers_rune_table) myers_rows = (int)std::min<uint32_t>(cs.longest, kRuneKernelLimit) + 1; // distinct runes <= length
    else if (myers_shared_table) {
        myers_rows = 1;
        for (int b = 0; b < 256; ++b) myers_rows += present[b] != 0;
    }
    if (unit_costs && engine.bit_parallel) {
        std::vector<uint32_t> kept;
        // Shared memory bounds the words per lane: a sixteenth of the SM per warp-private table (sixteen resident
        // warps per CTA), or the whole carve-out for a CTA-shared one.
        size_t const table_budget = myers_shared_table ? 220 * 1024 - 1024 : 12 * 1024;
        // Rune-keyed tables take one pattern word per lane: candidates of up to 1024 runes, longer ones stay on the DP tier.
        int const max_words = myers_rune_table ? 1 : (int)(table_budget / ((size_t)myers_rows * 32 * 4));
        uint32_t remembered_length = 0xFFFFFFFFu;
        int remembered_words = 0, remembered_lanes = 0;
        MyersLaunch *remembered_launch = nullptr;
        for (size_t i = n_candidates; i-- > 0;) { // longest first
            uint32_t const c = by_length[i], length = candidate_length(c);
            // Candidates arrive sorted by length: equal lengths are neighbours, so one remembered answer suffices.
            if (length != remembered_length) {
                myers_shape_for_length(length, max_words, (int)myers_min_lanes, remembered_words, remembered_lanes);
                remembered_length = length;
            }
            int const words = remembered_words, lanes = remembered_lanes;
            if (words == 0) {
                kept.push_back(c);
                continue;
            }
            if (!remembered_launch || remembered_launch->words != words) {
                auto launch = std::find_if(myers_launches.begin(), myers_launches.end(),
                                           [&](MyersLaunch const &l) { return l.words == words; });
                if (launch == myers_launches.end()) {
                    myers_launches.reserve(kMyersMaxWords); // keeps `remembered_launch` valid across push_backs
                    myers_launches.push_back({words, {}}), launch = myers_launches.end() - 1;
                }
                remembered_launch = &*launch;
            }
            remembered_launch->members.push_back({c, (uint32_t)lanes, rows_needed(c)});
            cells_myers += (uint64_t)length * query_bytes_in_rows;
        }
        std::reverse(kept.begin(), kept.end()); // back to ascending length for the slot builder
        by_length.swap(kept);
    }
    size_t const n_dp_candidates = by_length.size();
    // A launch that cannot fill the GPU for ~50 us is folded into the next wider one: padding a few candidates to
    // more words per lane costs less than a dependent launch of its own. (Measured and rejected: folding EVERY launch of
    // a small call into the widest one to save a ragged last round — c2 went from 37.8 to 35.8 TCUPS, i.e. exactly the
    // extra words; the SMs are ALU-bound, so idle warps of a ragged round cost nothing, padding does.)
    if (myers_launches.size() > 1) {
        std::sort(myers_launches.begin(), myers_launches.end(),
                  [](MyersLaunch const &a, MyersLaunch const &b) { return a.words < b.words; });
        for (size_t l = 0; l + 1 < myers_launches.size();) {
            uint64_t cells = 0;
            for (MyersCandidate const &member : myers_launches[l].members)
                cells += (uint64_t)candidate_length(member.candidate) * query_bytes_in_rows;
            if (cells >= 1500000000ull) {
                ++l;
                continue;
            }
            MyersLaunch &wider = myers_launches[l + 1];
            for (MyersCandidate member : myers_launches[l].members) {
                uint32_t lanes = 1;
                while (32u * wider.words * lanes < candidate_length(member.candidate)) lanes *= 2;
                member.lanes = std::max(lanes, member.lanes);
                wider.members.push_back(member);
            }
            myers_launches.erase(myers_launches.begin() + l);
        }
        for (MyersLaunch &launch : myers_launches)
            std::stable_sort(launch.members.begin(), launch.members.end(), [&](MyersCandidate const &a, MyersCandidate const &b) {
                return candidate_length(a.candidate) > candidate_length(b.candidate);
            });
    }
    // Heaviest launch first, so that the lighter ones fill the tail of the heavier ones.
    auto launch_weight = [&](MyersLaunch const &launch) {
        uint64_t weight = 0;
        for (MyersCandidate const &member : launch.members) weight += candidate_length(member.candidate);
        return weight;
    };
    std::sort(myers_launches.begin(), myers_launches.end(),
              [&](MyersLaunch const &a, MyersLaunch const &b) { return launch_weight(a) > launch_weight(b); });

    std::vector<HostSlot> &host_slots = schedule.slots;
    host_slots.clear();
    host_slots.reserve(n_dp_candidates);
    uint64_t &cells_total = schedule.cells_total = 0, &cells_u16 = schedule.cells_u16 = 0;
    // Strips wider than `widest_strip(family)` (launch.cuh) do not exist: such slots take equal passes of the widest one.
    auto shape_of = [&](int family, uint32_t longest) {
        Shape shape = shape_for_length(longest, latency_lanes >= 16 ? 32 : 8);
        int const cap = widest_strip(family);
        if (kStrips[shape.strip_index] > cap) {
            uint32_t const widest = (uint32_t)cap * 32;
            uint32_t const passes = (longest + widest - 1) / widest, width = (longest + passes - 1) / passes;
            for (int i = 0; i < kStripCount; ++i)
                if ((uint32_t)kStrips[i] * 32 >= width) return Shape {i, 32};
        }
        return shape;
    };
    for (size_t i = 0; i < n_dp_candidates;) {
        uint32_t const c0 = by_length[i], len0 = candidate_length(c0);
        bool const narrow0 = fits_u16(model, std::min<uint64_t>(len0, longest_query));
        if (narrow0 && i + 1 < n_dp_candidates) {
            uint32_t const c1 = by_length[i + 1], len1 = candidate_length(c1);
            if (fits_u16(model, std::min<uint64_t>(len1, longest_query))) {
                host_slots.push_back({{c0, c1, len0, len1}, model.family16, shape_of(model.family16, len1)});
                cells_u16 += ((uint64_t)len0 + len1) * query_bytes_in_rows;
                i += 2;
                continue;
            }
        }
        int const family = narrow0 ? model.family16 : model.family32;
        host_slots.push_back({{c0, kNoCandidate, len0, 0}, family, shape_of(family, len0)});
        if (narrow0) cells_u16 += (uint64_t)len0 * query_bytes_in_rows;
        i += 1;
    }
    for (size_t c = 0; c < n_candidates; ++c) cells_total += (uint64_t)candidate_length((uint32_t)c) * query_bytes_in_rows;

    // Every (family, strip width) is a persistent launch of its own with a ragged end of its own (launches cannot share
    // an SM): a launch with less than about a millisecond of work is folded into the next wider one of its family —
    // a few padded columns cost less than another tail. Matters for strong-scaled shards (c4 on 8 GPUs: 9 launches of
    // 1-3 ms each) much more than for whole tiles.
    {
        constexpr uint64_t min_launch_cells = 6000000000ull;
        for (int family = 0; family < kFamilyCount; ++family) {
            uint64_t cells[kStripCount] = {0};
            for (HostSlot const &host : host_slots)
                if (host.family == family)
                    cells[host.shape.strip_index] += ((uint64_t)host.slot.len0 + host.slot.len1) * query_bytes_in_rows;
            for (int strip = 0; strip + 1 < kStripCount; ++strip) {
                if (!cells[strip] || cells[strip] >= min_launch_cells) continue;
                int wider = -1;
                for (int candidate = strip + 1; candidate < kStripCount && wider < 0; ++candidate)
                    if (cells[candidate] && kStrips[candidate] <= widest_strip(family)) wider = candidate;
                if (wider < 0) continue;
                for (HostSlot &host : host_slots)
                    if (host.family == family && host.shape.strip_index == strip) host.shape.strip_index = wider;
                cells[wider] += cells[strip], cells[strip] = 0;
            }
        }
    }
    trace.mark("slots built");
    // 6. Jobs: every slot crossed with blocks of the query order; one launch per (family, strip width).
    std::vector<Job> &host_jobs = schedule.jobs;
    std::vector<LaunchPlan> &plans = schedule.plans;
    host_jobs.clear(), plans.clear();
    uint64_t &scratch_words_per_cta = schedule.scratch_words_per_cta = 0;
    for (int family = 0; family < kFamilyCount; ++family) {
        for (int strip = kStripCount - 1; strip >= 0; --strip) {
            // Longest slots first: their jobs are the heaviest, and the atomic job counter hands them out in order.
            std::vector<uint32_t> members;
            for (uint32_t s = (uint32_t)host_slots.size(); s-- > 0;)
                if (host_slots[s].family == family && host_slots[s].shape.strip_index == strip) members.push_back(s);
            if (members.empty()) continue;
            LaunchPlan plan = {family, strip, (uint32_t)host_jobs.size(), 0};
            bool const affine = family_is_affine(family);
            // The persistent CTAs drain the job list through one counter; the last job of every CTA ends at a different
            // time, so the tail of a launch idles about half a job per SM: jobs are kept to a few percent of a CTA's share.
            uint64_t const target_jobs = (uint64_t)device_sms * 24;
            uint64_t const wanted_blocks = (target_jobs + members.size() - 1) / members.size();
            for (uint32_t s : members) {
                CandSlot const &slot = host_slots[s].slot;
                Shape const shape = host_slots[s].shape;
                // A block is a whole number of rounds of the CTA's warps (each sweeps 32 / G queries at a time),
                // so that no warp idles at the job's barrier while others run an extra round.
                uint32_t const round = (uint32_t)(kThreads / 32) * (32 / shape.lanes);
                uint32_t const slot_rows =
                    std::max(rows_needed(slot.c0), slot.c1 != kNoCandidate ? rows_needed(slot.c1) : 0u);
                uint64_t const rounds_total = (slot_rows + round - 1) / round;
                uint64_t const blocks = std::max<uint64_t>(1, std::min<uint64_t>(wanted_blocks, rounds_total));
                uint32_t block_rows = (uint32_t)((rounds_total + blocks - 1) / blocks) * round;
                uint32_t const longest = std::max(slot.len0, slot.len1);
                bool const multi_pass = longest > (uint32_t)shape.coverage();
                if (multi_pass) {
                    // Boundary scratch of one job: rows of the block x DP rows of its longest query (x 2 for affine).
                    uint64_t const per_row = ((uint64_t)longest_query + 1) * (affine ? 2 : 1);
                    uint64_t const rows_cap = std::max<uint64_t>(round, kScratchWordsPerJob / per_row / round * round);
                    block_rows = (uint32_t)std::min<uint64_t>(block_rows, rows_cap);
                }
                for (uint32_t begin = 0; begin < slot_rows; begin += block_rows) {
                    uint32_t const end = (uint32_t)std::min<uint64_t>(slot_rows, (uint64_t)begin + block_rows);
                    if (multi_pass) {
                        uint64_t const rows = (uint64_t)query_length(order[begin]) + 1;
                        uint64_t const words = (uint64_t)(end - begin) * rows * (affine ? 2 : 1);
                        if (words > 0xFFFFFFFFull) return fail(sz_overflow_risk_k, "Boundary scratch too large");
                        scratch_words_per_cta = std::max<uint64_t>(scratch_words_per_cta, words);
                    }
                    host_jobs.push_back({s, begin, end, (uint32_t)shape.lanes});
                }
            }
            plan.n_jobs = (uint32_t)host_jobs.size() - plan.first_job;
            plans.push_back(plan);
        }
    }

    return ok();
}

/** Drains the scope's stream and the engine's side streams: after a failure nothing may still be running against
 *  buffers the next call reuses. */
static void drain(Engine &engine, cudaStream_t stream) {
    cudaStreamSynchronize(stream);
    for (cudaStream_t side : engine.side_streams)
        if (side) cudaStreamSynchronize(side);
    cudaGetLastError();
}

/** Inputs of at most this many bytes (both sides together) that the host can read are class-mapped on the host and
 *  travel inside the per-call arena: one host-to-device copy, no mapping kernels. */
constexpr uint64_t kPackedLimit = 256 * 1024;
/** Pageable result matrices up to this size are fetched through the engine's pinned buffer and scattered by the host. */
constexpr uint64_t kPinnedDownloadLimit = 8ull << 20;

/**
 *  @brief Scores query rows [row_begin, row_end) against every candidate and writes `results[q * stride + c]`.
 *         `candidates.object == nullptr` requests the symmetric self-similarity of the queries: the whole matrix
 *         (lower triangle scored, upper mirrored) when `whole_matrix`, else — a row shard — the cells
 *         (query i, candidate j) for i in the range and j < row_end, left for the caller to mirror once every shard is in.
 */
static Failure cross_unguarded(Engine &engine, Scope &scope, cudaStream_t stream, InputView queries_view,
                               InputView candidates_view, void *results, size_t stride, size_t row_begin, size_t row_end,
                               bool whole_matrix, size_t base_row) {
    Trace const trace;
    bool const symmetric = candidates_view.object == nullptr;
    szs_b200_call_stats_t stats = {};
    int const device_sms = std::max(1, device_facts(scope.gpu_device).sms);
    SZB_CUDA(cudaEventRecord(engine.events[0], stream));

    // 1. Read both layouts (host work; device-resident offsets are bounced).
    if (Failure failure = read_layout(queries_view, engine.queries, stream)) return failure;
    StagedSet &caller_queries = engine.queries;
    if (!symmetric)
        if (Failure failure = read_layout(candidates_view, engine.candidates, stream)) return failure;
    StagedSet &caller_candidates = symmetric ? engine.queries : engine.candidates;

    size_t const caller_rows = caller_queries.count, caller_columns = caller_candidates.count;
    if (whole_matrix) row_begin = 0, row_end = caller_rows;
    if (row_end > caller_rows || row_begin > row_end) return fail(sz_unexpected_dimensions_k, "Row range out of bounds");
    if (caller_columns && stride < caller_columns)
        return fail(sz_unexpected_dimensions_k, "Results row stride is smaller than the candidates count");
    if (row_end == row_begin || caller_columns == 0) {
        engine.stats = stats;
        return ok();
    }
    if (!results) return fail(sz_unexpected_dimensions_k, "Results pointer is null");
    if (caller_rows >= 0xFFFFFFF0ull || caller_columns >= 0xFFFFFFF0ull)
        return fail(sz_overflow_risk_k, "Too many strings for one call");

    // Small host-resident inputs never get a copy of their own: the host maps them to classes straight into the arena.
    // (Managed memory is left to the copy engine: touching it from the host would pull its pages back from the device.)
    bool packed = caller_queries.host_resident && caller_candidates.host_resident &&
                  caller_queries.total_bytes + (symmetric ? 0 : caller_candidates.total_bytes) <= kPackedLimit;
    if (!packed) {
        if (Failure failure = upload_raw(caller_queries, stream, stats)) return failure;
        if (!symmetric)
            if (Failure failure = upload_raw(caller_candidates, stream, stats)) return failure;
    }
    trace.mark("staged");

    // 2. Which symbols are we scoring: byte presence for the Levenshtein alphabet, UTF-8 decoding / rune ranking.
    uint32_t present[256] = {0};
    bool wide_symbols = false;
    if (Failure failure = resolve_symbols(engine, stream, symmetric, packed, device_sms, caller_queries,
                                          caller_candidates, stats, present, wide_symbols))
        return failure;
    trace.mark("alphabet known");
    Model model;
    build_model(engine, present, wide_symbols, model);
    if (!fits_s32(engine, (uint64_t)caller_queries.longest + caller_candidates.longest))
        return fail(sz_overflow_risk_k, "Pair too long for 32-bit DP cells");

    // Retrieval-shaped calls (a handful of queries against many candidates) trade roles: a kernel keeps the *candidate*
    // resident (match table / substitution profile) and streams 32 / G queries per warp sweep past it, so with fewer
    // queries than that most of each warp idles and every candidate pays its set-up for one string. Every score on this
    // path is symmetric under exchanging the two strings when the substitution costs are, so the few strings become the
    // resident side, the many stream, and the (columns x rows) result is transposed on the way home.
    bool const swapped = !symmetric && whole_matrix && model.symmetric_costs && caller_rows < kRoleSwapRows &&
                         caller_rows < caller_columns;
    StagedSet &qs = swapped ? caller_candidates : caller_queries;
    StagedSet &cs = swapped ? caller_queries : caller_candidates;
    size_t const n_queries = qs.count, n_candidates = cs.count;
    if (swapped) row_begin = 0, row_end = n_queries;
    size_t const n_rows = row_end - row_begin;
    // A row shard of a symmetric call scores its rows against candidates [0, row_end) only.
    size_t const n_columns = (symmetric && !whole_matrix) ? row_end : n_candidates;

    // 4-6. The schedule: query order, bit-parallel candidates, DP slots and shapes, jobs (pure host work).
    Schedule schedule;
    if (Failure failure = build_schedule(engine, model, qs, cs, symmetric, whole_matrix, row_begin, n_rows, n_columns,
                                         present, wide_symbols, device_sms, trace, schedule))
        return failure;
    std::vector<uint32_t> const &order = schedule.order, &position = schedule.position;
    bool const triangle = schedule.triangle;
    std::vector<HostSlot> const &host_slots = schedule.slots;
    std::vector<Job> const &host_jobs = schedule.jobs;
    std::vector<LaunchPlan> const &plans = schedule.plans;
    std::vector<MyersLaunch> const &myers_launches = schedule.myers_launches;
    bool const myers_shared_table = schedule.myers_shared_table, myers_rune_table = schedule.myers_rune_table;
    int const myers_rows = schedule.myers_rows;
    uint64_t const scratch_words_per_cta = schedule.scratch_words_per_cta;
    uint64_t const cells_total = schedule.cells_total, cells_u16 = schedule.cells_u16, cells_myers = schedule.cells_myers;
    auto query_length = [&](uint32_t q) { return qs.offsets[q + 1] - qs.offsets[q]; };
    trace.mark("jobs built");

    // 7. The arena: everything the host generated for this call, one host-to-device copy.
    size_t n_myers = 0;
    for (MyersLaunch const &launch : myers_launches) n_myers += launch.members.size();
    size_t const n_counters = plans.size() + myers_launches.size() + 16;
    bool const map_on_host = packed && !model.scoring.compare;
    ArenaLayout layout;
    size_t const at_q_offsets = layout.carve((n_queries + 1) * sizeof(uint32_t));
    size_t const at_c_offsets = symmetric ? at_q_offsets : layout.carve((n_candidates + 1) * sizeof(uint32_t));
    size_t const at_scoring = layout.carve(sizeof(Scoring));
    size_t const at_lut = layout.carve(256);
    size_t const at_order = layout.carve(order.size() * sizeof(uint32_t));
    size_t const at_position = layout.carve(position.size() * sizeof(uint32_t));
    size_t const at_slots = layout.carve(host_slots.size() * sizeof(CandSlot));
    size_t const at_jobs = layout.carve(host_jobs.size() * sizeof(Job));
    size_t const at_myers = layout.carve(n_myers * sizeof(MyersCandidate));
    size_t const at_counters = layout.carve(n_counters * sizeof(uint32_t));
    size_t const at_q_bytes = packed ? layout.carve(qs.total_bytes + 16) : 0;
    size_t const at_c_bytes = packed ? (symmetric ? at_q_bytes : layout.carve(cs.total_bytes + 16)) : 0;
    SZB_CUDA(engine.arena_host.reserve(layout.used));
    SZB_CUDA(engine.arena_device.reserve(layout.used));
    char *const host = (char *)engine.arena_host.ptr, *const device = (char *)engine.arena_device.ptr;
    {
        std::memcpy(host + at_q_offsets, qs.offsets.data(), (n_queries + 1) * sizeof(uint32_t));
        if (!symmetric) std::memcpy(host + at_c_offsets, cs.offsets.data(), (n_candidates + 1) * sizeof(uint32_t));
        std::memcpy(host + at_scoring, &model.scoring, sizeof(Scoring));
        std::memcpy(host + at_lut, model.lut, 256);
        std::memcpy(host + at_order, order.data(), order.size() * sizeof(uint32_t));
        if (!position.empty()) std::memcpy(host + at_position, position.data(), position.size() * sizeof(uint32_t));
        CandSlot *const slots_out = (CandSlot *)(host + at_slots);
        for (size_t s = 0; s < host_slots.size(); ++s) slots_out[s] = host_slots[s].slot;
        if (!host_jobs.empty()) std::memcpy(host + at_jobs, host_jobs.data(), host_jobs.size() * sizeof(Job));
        MyersCandidate *myers_out = (MyersCandidate *)(host + at_myers);
        for (MyersLaunch const &launch : myers_launches) {
            std::memcpy(myers_out, launch.members.data(), launch.members.size() * sizeof(MyersCandidate));
            myers_out += launch.members.size();
        }
        std::memset(host + at_counters, 0, n_counters * sizeof(uint32_t));
        if (packed) {
            auto pack = [&](StagedSet const &set, size_t at) {
                uint8_t *const out = (uint8_t *)(host + at);
                set.for_each_host_string([&](uint8_t const *start, size_t length, size_t i) {
                    uint8_t *const to = out + set.offsets[i];
                    if (map_on_host)
                        for (size_t k = 0; k < length; ++k) to[k] = model.lut[start[k]];
                    else if (length) std::memcpy(to, start, length);
                });
            };
            pack(qs, at_q_bytes);
            if (!symmetric) pack(cs, at_c_bytes);
        }
    }
    SZB_CUDA(cudaMemcpyAsync(device, host, layout.used, cudaMemcpyHostToDevice, stream));
    stats.h2d_bytes += layout.used;
    uint32_t const *q_offsets = (uint32_t const *)(device + at_q_offsets), *c_offsets = (uint32_t const *)(device + at_c_offsets);
    Scoring const *const device_scoring = (Scoring const *)(device + at_scoring);
    uint32_t const *const device_order = (uint32_t const *)(device + at_order);
    uint32_t *const device_counters = (uint32_t *)(device + at_counters);

    // 3. Bytes -> classes (skipped in compare mode, where the kernels read raw bytes or symbols).
    uint8_t const *q_classes, *c_classes;
    if (packed) q_classes = (uint8_t const *)(device + at_q_bytes), c_classes = (uint8_t const *)(device + at_c_bytes);
    else if (wide_symbols) q_classes = qs.runes.as<uint8_t>(), c_classes = cs.runes.as<uint8_t>();
    else if (model.scoring.compare) q_classes = qs.raw.as<uint8_t>(), c_classes = cs.raw.as<uint8_t>();
    else {
        auto map_set = [&](StagedSet &set) -> Failure {
            if (!set.total_bytes) return ok();
            uint64_t const vectors = (set.total_bytes + 15) / 16;
            SZB_LAUNCH(map_classes_kernel, helper_grid(vectors, 256, device_sms, 8), 256, 0, stream,
                       set.raw.as<uint8_t>(), set.classes.as<uint8_t>(), set.total_bytes,
                       (uint8_t const *)(device + at_lut));
            SZB_CUDA(cudaGetLastError());
            stats.kernel_launches++;
            return ok();
        };
        if (Failure failure = map_set(qs)) return failure;
        if (!symmetric)
            if (Failure failure = map_set(cs)) return failure;
        q_classes = qs.classes.as<uint8_t>(), c_classes = cs.classes.as<uint8_t>();
    }
    trace.mark("work list uploaded");

    SZB_CUDA(engine.results.reserve(n_rows * n_candidates * sizeof(int64_t)));
    if (swapped) SZB_CUDA(engine.results_home.reserve(n_rows * n_candidates * sizeof(int64_t)));
    // Launches overlap on side streams: every plan gets its own per-CTA boundary scratch, one slab per resident CTA.
    std::vector<uint64_t> scratch_base(plans.size() + 1, 0);
    for (size_t p = 0; p < plans.size(); ++p)
        scratch_base[p + 1] = scratch_base[p] + scratch_words_per_cta * std::min<uint64_t>(plans[p].n_jobs, (uint64_t)device_sms);
    if (scratch_words_per_cta) SZB_CUDA(engine.scratch.reserve(scratch_base.back() * 4));

    // 8. Launch: one persistent grid per (family, strip width) and per bit-parallel word count.
    SZB_CUDA(cudaEventRecord(engine.events[1], stream));
    size_t const n_launches = myers_launches.size() + plans.size();
    bool const fan_out = n_launches > 1;
    size_t launch_ordinal = 0;
    auto next_stream = [&]() -> cudaStream_t {
        size_t const slot = launch_ordinal++ % (Engine::kSideStreams + 1);
        return (fan_out && slot) ? engine.side_streams[slot - 1] : stream;
    };
    if (fan_out) {
        SZB_CUDA(cudaEventRecord(engine.fork_event, stream));
        for (cudaStream_t side : engine.side_streams) SZB_CUDA(cudaStreamWaitEvent(side, engine.fork_event, 0));
    }
    {
        int const warps = kThreads / 32;
        auto table_bytes = [&](MyersLaunch const &launch) -> size_t {
            return myers_rune_table     ? myers_rune_bytes(myers_rows)
                   : myers_shared_table ? myers_shared_bytes(launch.words, myers_rows)
                                        : (size_t)model.scoring.n_classes * 32 * launch.words * 4 * warps;
        };
        auto resident_ctas = [&](MyersLaunch const &launch) { return (!myers_shared_table && table_bytes(launch) * 2 <= 220 * 1024) ? 2 : 1; };
        // Several bit-parallel launches running side by side (one stream each) share the resident CTA slots in proportion
        // to their work — sweeps x steps x words — instead of each claiming all of them: see SZB_SHARE_RESIDENT_SLOTS.
        std::vector<uint64_t> slot_share(myers_launches.size(), 0);
        if (SZB_SHARE_RESIDENT_SLOTS && myers_launches.size() > 1 && n_launches <= (size_t)Engine::kSideStreams + 1) {
            int const per_sm = resident_ctas(myers_launches[0]);
            bool uniform = true;
            std::vector<long double> weight(myers_launches.size(), 0);
            long double total_weight = 0;
            for (size_t l = 0; l < myers_launches.size(); ++l) {
                uniform = uniform && resident_ctas(myers_launches[l]) == per_sm;
                for (MyersCandidate const &member : myers_launches[l].members)
                    weight[l] += (long double)member.rows * member.lanes * (myers_launches[l].words + 0.3L);
                total_weight += weight[l];
            }
            if (uniform && total_weight > 0) {
                uint64_t const slots = (uint64_t)device_sms * per_sm;
                uint64_t granted = 0;
                size_t heaviest = 0;
                for (size_t l = 0; l < myers_launches.size(); ++l) {
                    slot_share[l] = std::max<uint64_t>(1, (uint64_t)(slots * (weight[l] / total_weight)));
                    granted += slot_share[l];
                    if (weight[l] > weight[heaviest]) heaviest = l;
                }
                if (granted < slots) slot_share[heaviest] += slots - granted; // rounding leftovers
            }
        }
        size_t first = 0;
        for (size_t l = 0; l < myers_launches.size(); ++l) {
            MyersLaunch const &launch = myers_launches[l];
            size_t const smem = table_bytes(launch);
            int const ctas_per_sm = resident_ctas(launch);
            // Work units that pull tasks: warps (private tables) or whole CTAs (shared table).
            uint64_t const pullers = myers_shared_table ? (uint64_t)device_sms : (uint64_t)device_sms * ctas_per_sm * warps;
            uint32_t narrowest = 32;
            for (MyersCandidate const &member : launch.members) narrowest = std::min(narrowest, 32u / member.lanes);
            uint64_t const sweeps_per_candidate = (n_rows + narrowest - 1) / narrowest;
            uint64_t const total_sweeps = sweeps_per_candidate * launch.members.size();
            // ~16 tasks per puller; a shared-table task should also keep all sixteen warps busy for a few rounds.
            uint64_t batches_per_block = std::max<uint64_t>(1, total_sweeps / (pullers * 16));
            if (myers_shared_table) batches_per_block = std::max<uint64_t>(((batches_per_block + warps - 1) / warps) * warps, 2 * warps);
            uint64_t const blocks = (sweeps_per_candidate + batches_per_block - 1) / batches_per_block;
            if (blocks * launch.members.size() > 0xFFFFFFF0ull) return fail(sz_overflow_risk_k, "Too many tasks");
            MyersArgs args;
            args.q_cls = q_classes, args.q_off = q_offsets;
            args.c_cls = c_classes, args.c_off = c_offsets;
            args.q_order = device_order;
            args.candidates = (MyersCandidate const *)(device + at_myers) + first;
            args.n_candidates = (uint32_t)launch.members.size();
            args.n_rows = (uint32_t)n_rows;
            args.blocks_per_candidate = (uint32_t)blocks;
            args.batches_per_block = (uint32_t)batches_per_block;
            args.task_counter = device_counters + plans.size() + l;
            args.n_classes = model.scoring.n_classes;
            args.results = engine.results.as<int64_t>();
            args.stride = n_candidates;
            args.row_base = (uint32_t)row_begin;
            uint64_t const tasks = blocks * launch.members.size();
            uint64_t const resident = slot_share[l] ? slot_share[l] : (uint64_t)device_sms * ctas_per_sm;
            int const grid = myers_shared_table ? (int)std::min<uint64_t>(tasks, resident)
                                                : (int)std::min<uint64_t>((tasks + warps - 1) / warps, resident);
            int const table = myers_rune_table ? kMyersRuneTable : myers_shared_table ? kMyersSharedTable : kMyersPrivateTables;
            SZB_CUDA(launch_myers(launch.words, table, grid, kThreads, smem, next_stream(), args));
            stats.kernel_launches++;
            first += launch.members.size();
        }
    }
    for (size_t p = 0; p < plans.size(); ++p) {
        LaunchPlan const &plan = plans[p];
        KernelArgs args;
        args.q_cls = q_classes, args.q_off = q_offsets;
        args.c_cls = c_classes, args.c_off = c_offsets;
        args.wide_symbols = wide_symbols ? 1u : 0u;
        args.q_order = device_order;
        args.slots = (CandSlot const *)(device + at_slots);
        args.jobs = (Job const *)(device + at_jobs) + plan.first_job;
        args.n_jobs = plan.n_jobs;
        args.job_counter = device_counters + p;
        args.scoring = device_scoring;
        args.results = engine.results.as<int64_t>();
        args.stride = n_candidates;
        args.row_base = (uint32_t)row_begin;
        args.scratch = engine.scratch.as<uint32_t>() + scratch_base[p];
        args.scratch_words_per_cta = scratch_words_per_cta;
        int const grid = (int)std::min<uint32_t>(plan.n_jobs, (uint32_t)device_sms);
        size_t const smem = shared_bytes(plan.strip_index, model.scoring.n_classes, model.scoring.compare != 0);
        SZB_CUDA(launch_wavefront(plan.family, plan.strip_index, grid, smem, next_stream(), args));
        stats.kernel_launches++;
    }
    if (fan_out)
        for (int i = 0; i < Engine::kSideStreams; ++i) {
            SZB_CUDA(cudaEventRecord(engine.join_events[i], engine.side_streams[i]));
            SZB_CUDA(cudaStreamWaitEvent(stream, engine.join_events[i], 0));
        }
    if (symmetric && whole_matrix && n_queries > 1) {
        uint64_t const total = (uint64_t)n_queries * n_queries;
        int const blocks = helper_grid(total, 256, device_sms, 16);
        if (triangle)
            SZB_LAUNCH(mirror_ranked_kernel, blocks, 256, 0, stream, engine.results.as<int64_t>(), n_queries,
                       n_candidates, (uint32_t const *)(device + at_position));
        else
            SZB_LAUNCH(mirror_lower_kernel, blocks, 256, 0, stream, engine.results.as<int64_t>(), n_queries,
                       n_candidates);
        SZB_CUDA(cudaGetLastError());
        stats.kernel_launches++;
    }
    int64_t const *home = engine.results.as<int64_t>();
    if (swapped) {
        uint64_t const total = (uint64_t)caller_rows * caller_columns;
        SZB_LAUNCH(transpose_results_kernel, helper_grid(total, 256, device_sms, 16), 256, 0, stream,
                   engine.results.as<int64_t>(), engine.results_home.as<int64_t>(), caller_rows, caller_columns);
        SZB_CUDA(cudaGetLastError());
        stats.kernel_launches++;
        home = engine.results_home.as<int64_t>();
    }
    SZB_CUDA(cudaEventRecord(engine.events[2], stream));
    trace.mark("kernels launched");

    // 9. Results home: a strided 2-D copy into the caller's matrix (host, pinned, managed or device memory). Small
    //    pageable destinations are served through the engine's pinned buffer: one true DMA, then a host scatter.
    size_t const home_rows = swapped ? caller_rows : n_rows, home_columns = swapped ? caller_columns : n_columns;
    size_t const home_pitch = (swapped ? caller_columns : n_candidates) * sizeof(int64_t);
    int64_t *const destination = (int64_t *)results + (swapped ? 0 : row_begin - base_row) * stride;
    MemoryKind const results_kind = memory_kind(results);
    uint64_t const home_bytes = (uint64_t)home_rows * home_columns * sizeof(int64_t);
    bool const through_pinned = results_kind == MemoryKind::kPageable && home_bytes <= kPinnedDownloadLimit;
    if (through_pinned) {
        SZB_CUDA(engine.download.reserve(home_bytes));
        SZB_CUDA(copy_rows(engine.download.ptr, home_columns * sizeof(int64_t), home, home_pitch,
                           home_columns * sizeof(int64_t), home_rows, cudaMemcpyDeviceToHost, stream));
    }
    else {
        SZB_CUDA(copy_rows(destination, stride * sizeof(int64_t), home, home_pitch, home_columns * sizeof(int64_t),
                           home_rows, cudaMemcpyDefault, stream));
    }
    if (results_kind != MemoryKind::kDevice) stats.d2h_bytes += home_bytes;
    SZB_CUDA(cudaEventRecord(engine.events[3], stream));
    SZB_CUDA(cudaStreamSynchronize(stream));
    if (through_pinned)
        for (size_t r = 0; r < home_rows; ++r)
            std::memcpy(destination + r * stride, (int64_t const *)engine.download.ptr + r * home_columns,
                        home_columns * sizeof(int64_t));

    trace.mark("stream drained");
    float kernel_ms = 0, total_ms = 0;
    cudaEventElapsedTime(&kernel_ms, engine.events[1], engine.events[2]);
    cudaEventElapsedTime(&total_ms, engine.events[0], engine.events[3]);
    stats.kernel_milliseconds = kernel_ms, stats.total_milliseconds = total_ms;
    stats.cells = cells_total;
    if (symmetric && whole_matrix) {
        // Only the lower triangle (with the diagonal) counts as scored work, as in bench/similarities.cuh:344-363.
        uint64_t squares = 0;
        for (size_t i = 0; i < n_queries; ++i) squares += (uint64_t)query_length((uint32_t)i) * query_length((uint32_t)i);
        stats.cells = (cells_total + squares) / 2;
    }
    stats.pairs_u16 = cells_u16, stats.pairs_s32 = cells_total - cells_u16 - cells_myers, stats.pairs_bit_parallel = cells_myers;
    engine.stats = stats;
    return ok();
}

/** @param base_row the matrix row `results` points at: 0 for a full-shape matrix, `row_begin` for a compact slab. */
static Failure cross(Engine &engine, Scope &scope, InputView queries_view, InputView candidates_view, void *results,
                     size_t stride, size_t row_begin, size_t row_end, bool whole_matrix, size_t base_row = 0) {
    if (scope.kind == Scope::kCpu)
        return fail(sz_device_code_mismatch_k, "GPU engine cannot run on a CPU device scope");
    if (device_sm_code(scope.gpu_device) == 0 && visible_device_count() <= scope.gpu_device)
        return fail(sz_missing_gpu_k, "No CUDA device is visible");
    if (device_sm_code(scope.gpu_device) < 100)
        return fail(sz_missing_gpu_k, "stringzillas_b200 kernels require an sm_100-class (Blackwell) device");
    if (Failure failure = scope.ensure_stream()) return failure;
    std::lock_guard<std::mutex> engine_lock(engine.mutex);
    if (Failure failure = ensure_events(engine, scope.gpu_device)) return failure;
    Failure failure = cross_unguarded(engine, scope, scope.stream, queries_view, candidates_view, results, stride,
                                      row_begin, row_end, whole_matrix, whole_matrix ? 0 : base_row);
    if (failure) {
        // Launches of this call may still be in flight on the scope's stream or the side streams.
        drain(engine, scope.stream);
        engine.stats = {};
    }
    return failure;
}

#pragma endregion

} // namespace szb

namespace szb {

cudaError_t launch_wavefront(int family, int shape, int grid, size_t smem, cudaStream_t stream,
                             KernelArgs const &args) {
    switch (family) {
    case FAM_LIN_SHIFT: return launch_family<FAM_LIN_SHIFT>(shape, grid, smem, stream, args);
    case FAM_AFF_SHIFT: return launch_family<FAM_AFF_SHIFT>(shape, grid, smem, stream, args);
    case FAM_SW_LIN: return launch_family<FAM_SW_LIN>(shape, grid, smem, stream, args);
    case FAM_SW_AFF: return launch_family<FAM_SW_AFF>(shape, grid, smem, stream, args);
    case FAM_S32_GLOBAL_LIN: return launch_family<FAM_S32_GLOBAL_LIN>(shape, grid, smem, stream, args);
    case FAM_S32_GLOBAL_AFF: return launch_family<FAM_S32_GLOBAL_AFF>(shape, grid, smem, stream, args);
    case FAM_S32_LOCAL_LIN: return launch_family<FAM_S32_LOCAL_LIN>(shape, grid, smem, stream, args);
    case FAM_S32_LOCAL_AFF: return launch_family<FAM_S32_LOCAL_AFF>(shape, grid, smem, stream, args);
    default: return launch_family<FAM_LIN_COMPARE>(shape, grid, smem, stream, args);
    }
}

} // namespace szb

#pragma region Multi-GPU

namespace szb {

/**
 *  Contiguous blocks of query rows of near-equal cost: a row costs `len(q) * sum(len(c))`, so blocks
 *  balance the total query length; for a symmetric call only the cells left of the diagonal are scored, so row i weighs
 *  `len(i) * sum(len(j), j <= i)`. Every row weighs at least 1 so that empty strings still spread over the parts.
 *  Writes `parts + 1` bounds; deterministic, so every rank of a communicator computes the same partition.
 */
static void partition_rows(uint64_t const *lengths, size_t count, size_t parts, bool symmetric, uint64_t *bounds) {
    std::vector<long double> cumulative(count + 1, 0);
    long double prefix = 0;
    for (size_t i = 0; i < count; ++i) {
        long double const length = (long double)std::max<uint64_t>(lengths[i], 1);
        prefix += length;
        cumulative[i + 1] = cumulative[i] + (symmetric ? length * prefix : length);
    }
    bounds[0] = 0;
    for (size_t part = 1; part < parts; ++part) {
        long double const target = cumulative[count] * (long double)part / (long double)parts;
        size_t const cut = (size_t)(std::lower_bound(cumulative.begin(), cumulative.end(), target) - cumulative.begin());
        bounds[part] = std::min<uint64_t>(std::max<uint64_t>(cut, bounds[part - 1]), count);
    }
    bounds[parts] = count;
}

/** Host-readable copy of a u64 tape's string lengths (device-resident offsets are fetched once). */
static Failure tape_lengths(sz_sequence_u64tape_t const *tape, std::vector<uint64_t> &lengths) {
    size_t const count = (size_t)tape->count;
    lengths.assign(count, 0);
    if (!count) return ok();
    if (!tape->offsets) return fail(sz_unexpected_dimensions_k, "Tape offsets are null");
    std::vector<uint64_t> bounce;
    uint64_t const *offsets = (uint64_t const *)tape->offsets;
    if (!pointer_is_host_readable(offsets)) {
        bounce.resize(count + 1);
        SZB_CUDA(cudaMemcpy(bounce.data(), offsets, (count + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        offsets = bounce.data();
    }
    for (size_t i = 0; i < count; ++i) {
        if (offsets[i + 1] < offsets[i]) return fail(sz_unexpected_dimensions_k, "Tape offsets decrease");
        lengths[i] = offsets[i + 1] - offsets[i];
    }
    return ok();
}

/** Mirrors the lower triangle of a host-resident square matrix into the upper one, rows [begin, end), tile by tile. */
static void mirror_lower_on_host(int64_t *matrix, size_t n, size_t stride, size_t begin, size_t end) {
    constexpr size_t tile = 32;
    for (size_t i0 = begin; i0 < end; i0 += tile)
        for (size_t j0 = i0; j0 < n; j0 += tile)
            for (size_t i = i0; i < std::min(i0 + tile, end); ++i)
                for (size_t j = std::max(j0, i + 1); j < std::min(j0 + tile, n); ++j) matrix[i * stride + j] = matrix[j * stride + i];
}

/**
 *  The GPUs of one box driven together from ONE process: a scope (device + stream) and a parked host thread per GPU.
 *  The reference's scope constructor takes a single device id (stringzillas.h:137-138) and allows different scopes to
 *  be used from different host threads (types.cuh:433-444); this is the orchestrator SURVEY.md §8e asks for on top.
 */
struct DeviceGroup {
    std::vector<std::unique_ptr<Scope>> scopes;
    std::vector<std::thread> workers;
    std::mutex mutex;
    std::condition_variable wake, finished;
    uint64_t generation = 0;
    size_t pending = 0;
    bool stopping = false;
    void (*task)(void *, size_t) = nullptr;
    void *task_context = nullptr;
    std::mutex call_mutex; // one sharded call at a time per group

    ~DeviceGroup() {
        {
            std::lock_guard<std::mutex> lock(mutex);
            stopping = true;
        }
        wake.notify_all();
        for (std::thread &worker : workers) worker.join();
    }
    void start() {
        for (size_t index = 1; index < scopes.size(); ++index)
            workers.emplace_back([this, index] {
                uint64_t seen = 0;
                for (;;) {
                    void (*run)(void *, size_t);
                    void *context;
                    {
                        std::unique_lock<std::mutex> lock(mutex);
                        wake.wait(lock, [&] { return stopping || generation != seen; });
                        if (stopping) return;
                        seen = generation, run = task, context = task_context;
                    }
                    run(context, index);
                    {
                        std::lock_guard<std::mutex> lock(mutex);
                        --pending;
                    }
                    finished.notify_one();
                }
            });
    }
    /** Runs `run(context, i)` for every member: member 0 on the calling thread, the others on their parked threads. */
    void run_on_all(void (*run)(void *, size_t), void *context) {
        {
            std::lock_guard<std::mutex> lock(mutex);
            task = run, task_context = context, pending = workers.size(), ++generation;
        }
        wake.notify_all();
        run(context, 0);
        std::unique_lock<std::mutex> lock(mutex);
        finished.wait(lock, [&] { return pending == 0; });
    }
};

struct ShardedCall {
    Engine *engine;
    DeviceGroup *group;
    sz_sequence_u64tape_t const *queries, *candidates;
    void *results;
    size_t stride;
    std::vector<uint64_t> bounds;
    std::vector<Failure> failures;
    bool mirror_on_host = false;
    size_t n = 0;
};

static void run_shard(void *context, size_t index) {
    ShardedCall &call = *static_cast<ShardedCall *>(context);
    Engine &engine = index == 0 ? *call.engine : *call.engine->shards[index - 1];
    InputView q {InputView::kTape64, call.queries}, c {InputView::kTape64, call.candidates};
    size_t const begin = (size_t)call.bounds[index], end = (size_t)call.bounds[index + 1];
    bool const whole = call.group->scopes.size() == 1;
    call.failures[index] = cross(engine, *call.group->scopes[index], q, c, call.results, call.stride, begin, end, whole);
}

static void mirror_shard(void *context, size_t index) {
    ShardedCall &call = *static_cast<ShardedCall *>(context);
    mirror_lower_on_host((int64_t *)call.results, call.n, call.stride, (size_t)call.bounds[index], (size_t)call.bounds[index + 1]);
}

/**
 *  One cross product over all the GPUs of a group: rows are partitioned by cost, every shard copies the inputs it needs
 *  straight from the caller's memory and writes its rows straight into the caller's matrix. No collective is involved:
 *  host-originated data reaches each GPU over its own PCIe link.
 */
static Failure sharded_cross(Engine &engine, DeviceGroup &group, sz_sequence_u64tape_t const *queries,
                             sz_sequence_u64tape_t const *candidates, void *results, size_t stride) {
    std::lock_guard<std::mutex> call_lock(group.call_mutex);
    size_t const parts = group.scopes.size();
    while (engine.shards.size() + 1 < parts) {
        std::unique_ptr<Engine> shard = engine.clone();
        if (!shard) return fail(sz_bad_alloc_k, "Failed to allocate an engine shard");
        engine.shards.push_back(std::move(shard));
    }
    for (size_t i = 0; i + 1 < parts; ++i) engine.shards[i]->bit_parallel = engine.bit_parallel;
    ShardedCall call;
    call.engine = &engine, call.group = &group, call.queries = queries, call.candidates = candidates;
    call.results = results, call.stride = stride, call.n = (size_t)queries->count;
    std::vector<uint64_t> lengths;
    if (Failure failure = tape_lengths(queries, lengths)) return failure;
    bool const symmetric = candidates == nullptr;
    call.bounds.resize(parts + 1);
    partition_rows(lengths.data(), lengths.size(), parts, symmetric, call.bounds.data());
    call.failures.assign(parts, Failure {});
    auto const start = std::chrono::steady_clock::now();
    group.run_on_all(&run_shard, &call);
    for (Failure const &failure : call.failures)
        if (failure) return failure;
    // Row shards of a symmetric call leave the cells right of the diagonal to this mirror (serial.hpp:3183-3184).
    if (symmetric && parts > 1 && call.n > 1) {
        if (pointer_is_host_readable(results)) group.run_on_all(&mirror_shard, &call);
        else {
            Scope &scope = *group.scopes[0];
            SZB_CUDA(cudaSetDevice(scope.gpu_device));
            int const sms = std::max(1, device_facts(scope.gpu_device).sms);
            SZB_LAUNCH(mirror_lower_kernel, helper_grid((uint64_t)call.n * call.n, 256, sms, 16), 256, 0, scope.stream,
                       (int64_t *)results, call.n, stride);
            SZB_CUDA(cudaGetLastError());
            SZB_CUDA(cudaStreamSynchronize(scope.stream));
        }
    }
    double const wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start).count();
    // Telemetry of the whole call: the slowest shard's kernels, everything summed.
    szs_b200_call_stats_t total = engine.stats;
    for (size_t i = 0; i + 1 < parts; ++i) {
        szs_b200_call_stats_t const &shard = engine.shards[i]->stats;
        total.kernel_milliseconds = std::max(total.kernel_milliseconds, shard.kernel_milliseconds);
        total.cells += shard.cells, total.kernel_launches += shard.kernel_launches;
        total.h2d_bytes += shard.h2d_bytes, total.d2h_bytes += shard.d2h_bytes;
        total.pairs_u16 += shard.pairs_u16, total.pairs_s32 += shard.pairs_s32, total.pairs_bit_parallel += shard.pairs_bit_parallel;
    }
    if (parts > 1) total.total_milliseconds = wall_ms;
    engine.stats = total;
    return ok();
}

/* ---- One process per GPU: the exchange step over NCCL, issued from C ------------------------------------------- */

/** The handful of NCCL entry points the exchange needs, resolved at run time from the NCCL already in the process
 *  (torch's, under torch.distributed.run) or from `SZS_B200_NCCL_LIBRARY` / the system's libnccl.so.2. The library
 *  must load and score on one GPU without NCCL installed, so nothing here is a link-time dependency. */
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(void const *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(void const *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    char const *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ready() const { return GroupEnd != nullptr; }
};

static NcclApi const &nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        char const *override_path = std::getenv("SZS_B200_NCCL_LIBRARY");
        void *handle = override_path ? dlopen(override_path, RTLD_NOW | RTLD_GLOBAL) : nullptr;
        if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); // the copy torch brought along
        if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) return;
        NcclApi loaded;
        loaded.handle = handle;
        loaded.GetUniqueId = (decltype(loaded.GetUniqueId))dlsym(handle, "ncclGetUniqueId");
        loaded.CommInitRank = (decltype(loaded.CommInitRank))dlsym(handle, "ncclCommInitRank");
        loaded.CommDestroy = (decltype(loaded.CommDestroy))dlsym(handle, "ncclCommDestroy");
        loaded.Broadcast = (decltype(loaded.Broadcast))dlsym(handle, "ncclBroadcast");
        loaded.Send = (decltype(loaded.Send))dlsym(handle, "ncclSend");
        loaded.Recv = (decltype(loaded.Recv))dlsym(handle, "ncclRecv");
        loaded.GroupStart = (decltype(loaded.GroupStart))dlsym(handle, "ncclGroupStart");
        loaded.GetErrorString = (decltype(loaded.GetErrorString))dlsym(handle, "ncclGetErrorString");
        loaded.GroupEnd = (decltype(loaded.GroupEnd))dlsym(handle, "ncclGroupEnd");
        if (loaded.GetUniqueId && loaded.CommInitRank && loaded.CommDestroy && loaded.Broadcast && loaded.Send &&
            loaded.Recv && loaded.GroupStart && loaded.GroupEnd && loaded.GetErrorString)
            api = loaded;
    });
    return api;
}

#define SZB_NCCL(call)                                                       \
    do {                                                                     \
        ncclResult_t result_ = (call);                                       \
        if (result_ != ncclSuccess) return fail(sz_status_unknown_k, nccl_api().GetErrorString(result_)); \
    } while (0)

/** One rank of a row-sharded job: an NCCL communicator over the ranks' GPUs plus the exchange buffers. */
struct Communicator {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0, device = 0;
    DeviceBuffer header, payload, gathered;
    PinnedBuffer header_host, payload_host;
    std::vector<uint64_t> q_offsets, c_offsets, bounds;
    ~Communicator() {
        if (comm && nccl_api().ready()) nccl_api().CommDestroy(comm);
    }
};

constexpr size_t kHeaderWords = 8;
constexpr uint64_t kHeaderMagic = 0x30303242535A53ull; // "SZSB200"

/**
 *  @brief One cross product over the ranks of a communicator. Rank 0 owns the inputs (host or device memory) and
 *         receives the whole score matrix; the other ranks pass null tapes and results.
 *
 *  The only collectives are ONE broadcast of the packed inputs (a 64-byte header first, so that the other ranks can
 *  size the payload) and ONE gather of the score slabs (grouped send/recv, no padding), both on the scope's stream;
 *  every pair is independent, so nothing crosses GPUs during the DP itself.
 */
static Failure communicator_cross(Engine &engine, Scope &scope, Communicator &communicator,
                                  sz_sequence_u64tape_t const *queries, sz_sequence_u64tape_t const *candidates,
                                  void *results, size_t stride) {
    NcclApi const &nccl = nccl_api();
    if (!nccl.ready() || !communicator.comm) return fail(sz_status_unknown_k, "NCCL communicator is not initialised");
    if (scope.kind == Scope::kCpu) return fail(sz_device_code_mismatch_k, "GPU engine cannot run on a CPU device scope");
    if (Failure failure = scope.ensure_stream()) return failure;
    cudaStream_t const stream = scope.stream;
    Trace trace; // SZB200_TRACE=1: host-side marks of the exchange on rank 0 (with a stream sync before each mark)
    trace.enabled = trace.enabled && communicator.rank == 0;
    auto mark = [&](char const *what) {
        if (!trace.enabled) return;
        cudaStreamSynchronize(stream);
        trace.mark(what);
    };
    bool const root = communicator.rank == 0;
    int const world = communicator.world;
    SZB_CUDA(communicator.header.reserve(kHeaderWords * sizeof(uint64_t)));
    SZB_CUDA(communicator.header_host.reserve(kHeaderWords * sizeof(uint64_t)));
    uint64_t *const header = (uint64_t *)communicator.header_host.ptr;

    // 1. Rank 0 packs both tapes into one message: offsets (relative, 64-bit) first, then the bytes.
    uint64_t n_q = 0, n_c = 0, bytes_q = 0, bytes_c = 0, symmetric = 0, q_first = 0, c_first = 0;
    // Whatever rank 0 finds wrong with its arguments travels in the header: the other ranks are already waiting in the
    // broadcast, and a rank that returned early would leave them there.
    Failure rejected;
    if (root) {
        if (!queries) rejected = fail(sz_status_unknown_k, "Rank 0 must pass the queries");
        symmetric = candidates == nullptr;
        auto measure = [](sz_sequence_u64tape_t const *tape, std::vector<uint64_t> &relative, uint64_t &bytes,
                          uint64_t &first) -> Failure {
            size_t const count = (size_t)tape->count;
            relative.assign(count + 1, 0);
            bytes = 0, first = 0;
            if (!count) return ok();
            if (!tape->offsets) return fail(sz_unexpected_dimensions_k, "Tape offsets are null");
            std::vector<uint64_t> bounce;
            uint64_t const *offsets = (uint64_t const *)tape->offsets;
            if (!pointer_is_host_readable(offsets)) {
                bounce.resize(count + 1);
                SZB_CUDA(cudaMemcpy(bounce.data(), offsets, (count + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
                offsets = bounce.data();
            }
            for (size_t i = 0; i <= count; ++i) {
                if (i && offsets[i] < offsets[i - 1]) return fail(sz_unexpected_dimensions_k, "Tape offsets decrease");
                relative[i] = offsets[i] - offsets[0];
            }
            bytes = relative[count], first = offsets[0];
            return ok();
        };
        if (!rejected) {
            n_q = queries->count;
            rejected = measure(queries, communicator.q_offsets, bytes_q, q_first);
        }
        if (!rejected && !symmetric) {
            n_c = candidates->count;
            rejected = measure(candidates, communicator.c_offsets, bytes_c, c_first);
        }
        if (!rejected && n_q && (symmetric ? n_q : n_c)) {
            if (!results) rejected = fail(sz_unexpected_dimensions_k, "Results pointer is null");
            else if (stride < (symmetric ? n_q : n_c))
                rejected = fail(sz_unexpected_dimensions_k, "Results row stride is smaller than the candidates count");
        }
        if (rejected) n_q = n_c = bytes_q = bytes_c = 0;
        header[0] = kHeaderMagic, header[1] = n_q, header[2] = n_c, header[3] = bytes_q, header[4] = bytes_c;
        header[5] = symmetric, header[6] = (uint64_t)(int64_t)rejected.status, header[7] = 0;
        SZB_CUDA(cudaMemcpyAsync(communicator.header.ptr, header, kHeaderWords * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
    }
    SZB_NCCL(nccl.Broadcast(communicator.header.ptr, communicator.header.ptr, kHeaderWords * sizeof(uint64_t), ncclUint8, 0,
                            communicator.comm, stream));
    if (!root) {
        SZB_CUDA(cudaMemcpyAsync(header, communicator.header.ptr, kHeaderWords * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
        SZB_CUDA(cudaStreamSynchronize(stream));
        if (header[0] != kHeaderMagic) return fail(sz_status_unknown_k, "Malformed exchange header");
        n_q = header[1], n_c = header[2], bytes_q = header[3], bytes_c = header[4], symmetric = header[5];
        if (header[6]) return fail((sz_status_t)(int64_t)header[6], "Rank 0 rejected its inputs");
    }
    else if (rejected) {
        SZB_CUDA(cudaStreamSynchronize(stream)); // the header has left before this rank does
        return rejected;
    }
    size_t const offsets_bytes = (size_t)((n_q + 1) + (symmetric ? 0 : n_c + 1)) * sizeof(uint64_t);
    size_t const at_q_bytes = (offsets_bytes + 15) & ~(size_t)15;
    size_t const at_c_bytes = (at_q_bytes + bytes_q + 15) & ~(size_t)15;
    size_t const payload_bytes = ((at_c_bytes + bytes_c + 15) & ~(size_t)15) + 16;
    SZB_CUDA(communicator.payload.reserve(payload_bytes));
    char *const payload = (char *)communicator.payload.ptr;
    if (root) {
        // Offsets go through a pinned buffer; the bytes are copied from wherever the caller keeps them.
        SZB_CUDA(communicator.payload_host.reserve(offsets_bytes));
        uint64_t *const packed = (uint64_t *)communicator.payload_host.ptr;
        std::memcpy(packed, communicator.q_offsets.data(), (n_q + 1) * sizeof(uint64_t));
        if (!symmetric) std::memcpy(packed + n_q + 1, communicator.c_offsets.data(), (n_c + 1) * sizeof(uint64_t));
        SZB_CUDA(cudaMemcpyAsync(payload, packed, offsets_bytes, cudaMemcpyHostToDevice, stream));
        if (bytes_q)
            SZB_CUDA(cudaMemcpyAsync(payload + at_q_bytes, (char const *)queries->data + q_first, bytes_q, cudaMemcpyDefault, stream));
        if (!symmetric && bytes_c) {
            SZB_CUDA(cudaMemcpyAsync(payload + at_c_bytes, (char const *)candidates->data + c_first, bytes_c, cudaMemcpyDefault, stream));
        }
    }
    mark("exchange: header broadcast and inputs packed");
    SZB_NCCL(nccl.Broadcast(payload, payload, payload_bytes, ncclUint8, 0, communicator.comm, stream));
    mark("exchange: payload broadcast");
    if (!root) {
        communicator.q_offsets.resize(n_q + 1);
        SZB_CUDA(cudaMemcpyAsync(communicator.q_offsets.data(), payload, (n_q + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
        if (!symmetric) {
            communicator.c_offsets.resize(n_c + 1);
            SZB_CUDA(cudaMemcpyAsync(communicator.c_offsets.data(), payload + (n_q + 1) * sizeof(uint64_t),
                                     (n_c + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
        }
        SZB_CUDA(cudaStreamSynchronize(stream));
    }

    // 2. Every rank scores its block of rows against the device-resident copy.
    size_t const columns = symmetric ? (size_t)n_q : (size_t)n_c;
    std::vector<uint64_t> lengths(n_q);
    for (size_t i = 0; i < n_q; ++i) lengths[i] = communicator.q_offsets[i + 1] - communicator.q_offsets[i];
    communicator.bounds.resize(world + 1);
    partition_rows(lengths.data(), (size_t)n_q, (size_t)world, symmetric != 0, communicator.bounds.data());
    size_t const begin = (size_t)communicator.bounds[communicator.rank], end = (size_t)communicator.bounds[communicator.rank + 1];
    sz_sequence_u64tape_t local_q {payload + at_q_bytes, communicator.q_offsets.data(), (sz_size_t)n_q};
    sz_sequence_u64tape_t local_c {payload + at_c_bytes, communicator.c_offsets.data(), (sz_size_t)n_c};
    InputView q {InputView::kTape64, &local_q}, c {InputView::kTape64, symmetric ? nullptr : &local_c};
    // Rank 0 assembles the whole matrix on its device; the others keep only their slab.
    size_t const slab_rows = root ? (size_t)n_q : end - begin;
    SZB_CUDA(communicator.gathered.reserve(std::max<size_t>(slab_rows * columns, 1) * sizeof(int64_t)));
    int64_t *const slab = communicator.gathered.as<int64_t>();
    bool const whole = world == 1;
    if (n_q && columns)
        if (Failure failure = cross(engine, scope, q, c, slab, columns, begin, end, whole, root ? 0 : begin)) return failure;

    mark("exchange: own rows scored");
    // 3. One gather of the slabs (rows x columns of int64, contiguous on both ends).
    if (world > 1 && columns) {
        SZB_NCCL(nccl.GroupStart());
        if (root) {
            for (int peer = 1; peer < world; ++peer) {
                size_t const peer_rows = (size_t)(communicator.bounds[peer + 1] - communicator.bounds[peer]);
                if (peer_rows)
                    SZB_NCCL(nccl.Recv(slab + (size_t)communicator.bounds[peer] * columns, peer_rows * columns * sizeof(int64_t),
                                       ncclUint8, peer, communicator.comm, stream));
            }
        }
        else if (end > begin)
            SZB_NCCL(nccl.Send(slab, (end - begin) * columns * sizeof(int64_t), ncclUint8, 0, communicator.comm, stream));
        SZB_NCCL(nccl.GroupEnd());
    }
    mark("exchange: slabs gathered");
    if (root && n_q && columns) {
        if (!results) return fail(sz_unexpected_dimensions_k, "Results pointer is null");
        if (stride < columns) return fail(sz_unexpected_dimensions_k, "Results row stride is smaller than the candidates count");
        if (symmetric && world > 1 && n_q > 1) {
            int const sms = std::max(1, device_facts(scope.gpu_device).sms);
            SZB_LAUNCH(mirror_lower_kernel, helper_grid(n_q * n_q, 256, sms, 16), 256, 0, stream, slab, n_q, columns);
            SZB_CUDA(cudaGetLastError());
        }
        SZB_CUDA(copy_rows(results, stride * sizeof(int64_t), slab, columns * sizeof(int64_t), columns * sizeof(int64_t),
                           (size_t)n_q, cudaMemcpyDefault, stream));
    }
    SZB_CUDA(cudaStreamSynchronize(stream));
    mark("exchange: matrix home");
    return ok();
}

} // namespace szb

#pragma endregion

#pragma region C ABI

using szb::Engine;
using szb::Failure;
using szb::InputView;
using szb::Scope;

extern "C" {

SZS_B200_API int szs_version_major(void) { return 5; } // tracks the reference ABI it drops in for: 5.1.2
SZS_B200_API int szs_version_minor(void) { return 1; }
SZS_B200_API int szs_version_patch(void) { return 2; }

SZS_B200_API sz_capability_t szs_capabilities_comptime(void) {
    return (sz_capability_t)(sz_cap_serial_k | sz_caps_cuda_k);
}

SZS_B200_API sz_capability_t szs_capabilities_runtime(void) {
    // Like the reference (runtime.cuh:27-45): an unreachable GPU contributes no bits instead of an error.
    int caps = sz_cap_serial_k;
    for (int device = 0, count = szb::visible_device_count(); device < count; ++device)
        if (szb::device_sm_code(device) >= 100) caps |= sz_caps_cuda_k;
    return (sz_capability_t)caps;
}

SZS_B200_API sz_capability_t szs_capabilities(void) {
    return (sz_capability_t)(szs_capabilities_comptime() & szs_capabilities_runtime());
}

static void *unified_allocate(sz_size_t bytes, void *) { return szs_unified_alloc(bytes); }
static void unified_free(void *pointer, sz_size_t bytes, void *) { szs_unified_free(pointer, bytes); }

SZS_B200_API sz_status_t sz_memory_allocator_init_unified(sz_memory_allocator_t *alloc, char const **error_message) {
    if (!alloc) return szb::report(szb::fail(sz_status_unknown_k, "Allocator pointer is null"), error_message);
    alloc->allocate = &unified_allocate, alloc->free = &unified_free, alloc->handle = nullptr;
    return szb::report(szb::ok(), error_message);
}

SZS_B200_API void *szs_unified_alloc(sz_size_t size_bytes) {
    void *pointer = nullptr;
    if (cudaMallocManaged(&pointer, size_bytes ? size_bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return pointer;
}

SZS_B200_API void szs_unified_free(void *ptr, sz_size_t) {
    if (ptr) cudaFree(ptr);
}

SZS_B200_API sz_status_t szs_device_scope_init_default(szs_device_scope_t *scope, char const **error_message) {
    if (!scope) return szb::report(szb::fail(sz_status_unknown_k, "Scope pointer is null"), error_message);
    auto *made = new (std::nothrow) Scope();
    if (!made) return szb::report(szb::fail(sz_bad_alloc_k, "Failed to allocate device scope"), error_message);
    made->kind = Scope::kDefault, made->cpu_cores = 1, made->gpu_device = 0;
    *scope = made;
    return szb::report(szb::ok(), error_message);
}

SZS_B200_API sz_status_t szs_device_scope_init_cpu_cores(sz_size_t cpu_cores, szs_device_scope_t *scope,
                                                         char const **error_message) {
    if (!scope) return szb::report(szb::fail(sz_status_unknown_k, "Scope pointer is null"), error_message);
    // A one-core request folds back onto the default scope, as in the reference (runtime.cuh:93-94).
    if (cpu_cores == 1) return szs_device_scope_init_default(scope, error_message);
    auto *made = new (std::nothrow) Scope();
    if (!made) return szb::report(szb::fail(sz_bad_alloc_k, "Failed to allocate device scope"), error_message);
    made->kind = Scope::kCpu, made->cpu_cores = cpu_cores, made->gpu_device = -1;
    *scope = made;
    return szb::report(szb::ok(), error_message);
}

SZS_B200_API sz_status_t szs_device_scope_init_gpu_device(sz_size_t gpu_device, szs_device_scope_t *scope,
                                                          char const **error_message) {
    if (!scope) return szb::report(szb::fail(sz_status_unknown_k, "Scope pointer is null"), error_message);
    if ((size_t)szb::visible_device_count() <= gpu_device)
        return szb::report(szb::fail(sz_missing_gpu_k, "No such CUDA device"), error_message);
    auto *made = new (std::nothrow) Scope();
    if (!made) return szb::report(szb::fail(sz_bad_alloc_k, "Failed to allocate device scope"), error_message);
    made->kind = Scope::kGpu, made->cpu_cores = 0, made->gpu_device = (int)gpu_device;
    *scope = made;
    return szb::report(szb::ok(), error_message);
}

SZS_B200_API sz_status_t szs_device_scope_get_cpu_cores(szs_device_scope_t scope, sz_size_t *cpu_cores,
                                                        char const **error_message) {
    if (!scope || !cpu_cores) return szb::report(szb::fail(sz_status_unknown_k, "Invalid null pointer argument"), error_message);
    auto *typed = static_cast<Scope *>(scope);
    if (typed->kind == Scope::kGpu) return szb::report(szb::fail(sz_status_unknown_k, "Device scope is GPU-only"), error_message);
    *cpu_cores = typed->cpu_cores;
    return szb::report(szb::ok(), error_message);
}

SZS_B200_API sz_status_t szs_device_scope_get_gpu_device(szs_device_scope_t scope, sz_size_t *gpu_device,
                                                         char const **error_message) {
    if (!scope || !gpu_device) return szb::report(szb::fail(sz_status_unknown_k, "Invalid null pointer argument"), error_message);
    auto *typed = static_cast<Scope *>(scope);
    if (typed->kind != Scope::kGpu) return szb::report(szb::fail(sz_status_unknown_k, "Device scope is CPU-only"), error_message);
    *gpu_device = (sz_size_t)typed->gpu_device;
    return szb::report(szb::ok(), error_message);
}

SZS_B200_API sz_status_t szs_device_scope_get_capabilities(szs_device_scope_t scope, sz_capability_t *capabilities,
                                                           char const **error_message) {
    if (!scope || !capabilities) return szb::report(szb::fail(sz_status_unknown_k, "Invalid null pointer argument"), error_message);
    auto *typed = static_cast<Scope *>(scope);
    int const system = szs_capabilities();
    *capabilities = (sz_capability_t)(typed->kind == Scope::kGpu ? (system & sz_caps_cuda_k) : (system & sz_cap_serial_k));
    return szb::report(szb::ok(), error_message);
}

SZS_B200_API void szs_device_scope_free(szs_device_scope_t scope) { delete static_cast<Scope *>(scope); }

/* ---- Engines ---------------------------------------------------------------------------------------------------- */

static sz_status_t make_engine(szb::Algorithm algorithm, void **handle, char const **error_message, Engine **made) {
    if (!handle) return szb::report(szb::fail(sz_status_unknown_k, "Engine pointer is null"), error_message);
    auto *engine = new (std::nothrow) Engine();
    if (!engine) return szb::report(szb::fail(sz_bad_alloc_k, "Failed to allocate the engine"), error_message);
    engine->algorithm = algorithm;
    *handle = engine, *made = engine;
    return szb::report(szb::ok(), error_message);
}

SZS_B200_API sz_status_t szs_levenshtein_distances_init(sz_error_cost_t match, sz_error_cost_t mismatch,
                                                        sz_error_cost_t open, sz_error_cost_t extend,
                                                        sz_memory_allocator_t const *, sz_capability_t,
                                                        szs_levenshtein_distances_t *engine,
                                                        char const **error_message) {
    Engine *made = nullptr;
    sz_status_t status = make_engine(szb::Algorithm::kLevenshtein, engine, error_message, &made);
    if (status != sz_success_k) return status;
    made->match = match, made->mismatch = mismatch, made->open = open, made->extend = extend;
    return status;
}

SZS_B200_API sz_status_t szs_levenshtein_distances_utf8_init(sz_error_cost_t match, sz_error_cost_t mismatch,
                                                             sz_error_cost_t open, sz_error_cost_t extend,
                                                             sz_memory_allocator_t const *, sz_capability_t,
                                                             szs_levenshtein_distances_utf8_t *engine,
                                                             char const **error_message) {
    Engine *made = nullptr;
    sz_status_t status = make_engine(szb::Algorithm::kLevenshtein, engine, error_message, &made);
    if (status != sz_success_k) return status;
    made->match = match, made->mismatch = mismatch, made->open = open, made->extend = extend;
    made->utf8 = true;
    return status;
}

static sz_status_t init_weighted(szb::Algorithm algorithm, sz_u8_t const *byte_to_class,
                                 sz_error_cost_t const *class_substitution_costs, sz_error_cost_t open,
                                 sz_error_cost_t extend, void **engine, char const **error_message) {
    if (!byte_to_class || !class_substitution_costs)
        return szb::report(szb::fail(sz_status_unknown_k, "Substitution tables are null"), error_message);
    // The reference indexes its 32x32 table with whatever the map holds (serial.hpp:196-203: out-of-range classes read
    // past the table). Refuse them here instead: the profile builder would otherwise do the same on device memory.
    for (int byte = 0; byte < 256; ++byte)
        if (byte_to_class[byte] >= 32)
            return szb::report(szb::fail(sz_unexpected_dimensions_k, "byte_to_class maps a byte to a class index above 31"),
                               error_message);
    Engine *made = nullptr;
    sz_status_t status = make_engine(algorithm, engine, error_message, &made);
    if (status != sz_success_k) return status;
    std::memcpy(made->byte_to_class, byte_to_class, 256);
    std::memcpy(made->costs, class_substitution_costs, 32 * 32);
    made->open = open, made->extend = extend;
    return status;
}

SZS_B200_API sz_status_t szs_needleman_wunsch_scores_init(sz_u8_t const *byte_to_class,
                                                          sz_error_cost_t const *class_substitution_costs,
                                                          sz_error_cost_t open, sz_error_cost_t extend,
                                                          sz_memory_allocator_t const *, sz_capability_t,
                                                          szs_needleman_wunsch_scores_t *engine,
                                                          char const **error_message) {
    return init_weighted(szb::Algorithm::kNeedlemanWunsch, byte_to_class, class_substitution_costs, open, extend,
                         engine, error_message);
}

SZS_B200_API sz_status_t szs_smith_waterman_scores_init(sz_u8_t const *byte_to_class,
                                                        sz_error_cost_t const *class_substitution_costs,
                                                        sz_error_cost_t open, sz_error_cost_t extend,
                                                        sz_memory_allocator_t const *, sz_capability_t,
                                                        szs_smith_waterman_scores_t *engine,
                                                        char const **error_message) {
    return init_weighted(szb::Algorithm::kSmithWaterman, byte_to_class, class_substitution_costs, open, extend, engine,
                         error_message);
}

static sz_status_t run(void *engine, szb::Algorithm expected, szs_device_scope_t device, InputView::Layout layout,
                       void const *queries, void const *candidates, void *results, sz_size_t stride,
                       char const **error_message) {
    if (!engine || !device || !queries)
        return szb::report(szb::fail(sz_status_unknown_k, "Engine, device scope and queries must not be null"), error_message);
    auto *typed = static_cast<Engine *>(engine);
    if (typed->algorithm != expected)
        return szb::report(szb::fail(sz_status_unknown_k, "Engine handle belongs to another algorithm"), error_message);
    InputView q {layout, queries}, c {layout, candidates};
    Failure failure = szb::cross(*typed, *static_cast<Scope *>(device), q, c, results, stride, 0, 0, true);
    return szb::report(failure, error_message);
}

#define SZB_DEFINE_CALLS(name, algorithm, result_type)                                                                 \
    SZS_B200_API sz_status_t name(void *engine, szs_device_scope_t device, sz_sequence_t const *queries,                \
                                  sz_sequence_t const *candidates, result_type *results, sz_size_t stride,             \
                                  char const **error_message) {                                                        \
        return run(engine, algorithm, device, InputView::kSequence, queries, candidates, results, stride,              \
                   error_message);                                                                                     \
    }                                                                                                                  \
    SZS_B200_API sz_status_t name##_u32tape(void *engine, szs_device_scope_t device,                                    \
                                            sz_sequence_u32tape_t const *queries,                                      \
                                            sz_sequence_u32tape_t const *candidates, result_type *results,             \
                                            sz_size_t stride, char const **error_message) {                            \
        return run(engine, algorithm, device, InputView::kTape32, queries, candidates, results, stride,                \
                   error_message);                                                                                     \
    }                                                                                                                  \
    SZS_B200_API sz_status_t name##_u64tape(void *engine, szs_device_scope_t device,                                    \
                                            sz_sequence_u64tape_t const *queries,                                      \
                                            sz_sequence_u64tape_t const *candidates, result_type *results,             \
                                            sz_size_t stride, char const **error_message) {                            \
        return run(engine, algorithm, device, InputView::kTape64, queries, candidates, results, stride,                \
                   error_message);                                                                                     \
    }                                                                                                                  \
    SZS_B200_API void name##_free(void *engine) { delete static_cast<Engine *>(engine); }

SZB_DEFINE_CALLS(szs_levenshtein_distances, szb::Algorithm::kLevenshtein, sz_size_t)
SZB_DEFINE_CALLS(szs_levenshtein_distances_utf8, szb::Algorithm::kLevenshtein, sz_size_t)
SZB_DEFINE_CALLS(szs_needleman_wunsch_scores, szb::Algorithm::kNeedlemanWunsch, sz_ssize_t)
SZB_DEFINE_CALLS(szs_smith_waterman_scores, szb::Algorithm::kSmithWaterman, sz_ssize_t)

/* ---- Out-of-scope family (fingerprints): present so that bindings resolve, refusing loudly (no CPU fallback). ---- */

static sz_status_t out_of_scope(char const **error_message) {
    return szb::report(szb::fail(sz_missing_gpu_k, "Not part of the stringzillas_b200 hot path"), error_message);
}

SZS_B200_API sz_status_t szs_fingerprints_init(sz_size_t, sz_size_t, sz_size_t const *, sz_size_t, sz_u64_t,
                                               sz_memory_allocator_t const *, sz_capability_t, szs_fingerprints_t *,
                                               char const **error_message) {
    return out_of_scope(error_message);
}
SZS_B200_API sz_status_t szs_fingerprints_sequence(szs_fingerprints_t, szs_device_scope_t, sz_sequence_t const *,
                                                   sz_u32_t *, sz_size_t, sz_u32_t *, sz_size_t,
                                                   char const **error_message) {
    return out_of_scope(error_message);
}
SZS_B200_API sz_status_t szs_fingerprints_u32tape(szs_fingerprints_t, szs_device_scope_t,
                                                  sz_sequence_u32tape_t const *, sz_u32_t *, sz_size_t, sz_u32_t *,
                                                  sz_size_t, char const **error_message) {
    return out_of_scope(error_message);
}
SZS_B200_API sz_status_t szs_fingerprints_u64tape(szs_fingerprints_t, szs_device_scope_t,
                                                  sz_sequence_u64tape_t const *, sz_u32_t *, sz_size_t, sz_u32_t *,
                                                  sz_size_t, char const **error_message) {
    return out_of_scope(error_message);
}
SZS_B200_API void szs_fingerprints_free(szs_fingerprints_t) {}

/* ---- B200 additions ---------------------------------------------------------------------------------------------- */

SZS_B200_API sz_status_t szs_b200_last_call_stats(void const *engine, szs_b200_call_stats_t *stats) {
    if (!engine || !stats) return sz_status_unknown_k;
    *stats = static_cast<Engine const *>(engine)->stats;
    return sz_success_k;
}

SZS_B200_API sz_status_t szs_b200_debug_schedule(void const *engine_punned, sz_u32_t const *query_lengths,
                                                 sz_size_t n_queries, sz_u32_t const *candidate_lengths,
                                                 sz_size_t n_candidates, int distinct_candidate_bytes, int device_sms,
                                                 sz_u32_t *assignment, sz_u32_t *partner, sz_u32_t *order,
                                                 sz_u32_t *jobs, sz_size_t jobs_capacity, sz_size_t *n_jobs) {
    if (!engine_punned || !query_lengths || !assignment || !partner || !order || !n_jobs) return sz_status_unknown_k;
    auto const &engine = *static_cast<Engine const *>(engine_punned);
    bool const symmetric = candidate_lengths == nullptr;
    szb::StagedSet qs, cs_storage;
    auto fill = [](szb::StagedSet &set, sz_u32_t const *lengths, size_t count) {
        set.count = count, set.offsets.assign(count + 1, 0), set.longest = 0;
        for (size_t i = 0; i < count; ++i)
            set.offsets[i + 1] = set.offsets[i] + lengths[i], set.longest = std::max(set.longest, lengths[i]);
        set.total_bytes = set.offsets[count];
    };
    fill(qs, query_lengths, n_queries);
    if (!symmetric) fill(cs_storage, candidate_lengths, n_candidates);
    szb::StagedSet const &cs = symmetric ? qs : cs_storage;
    if (qs.count == 0 || cs.count == 0) return sz_unexpected_dimensions_k;
    uint32_t present[256] = {0};
    for (int b = 0; b < 256 && b < distinct_candidate_bytes; ++b) present[b] = 1;
    szb::Model model;
    szb::build_model(engine, present, false, model);
    szb::Schedule schedule;
    szb::Trace const trace;
    szb::Failure failure = szb::build_schedule(engine, model, qs, cs, symmetric, true, 0, qs.count, cs.count, present,
                                               false, std::max(device_sms, 1), trace, schedule);
    if (failure) return failure.status;
    for (size_t c = 0; c < cs.count; ++c) assignment[c] = 0xFFFFFFFFu, partner[c] = 0xFFFFFFFFu;
    for (szb::MyersLaunch const &launch : schedule.myers_launches)
        for (szb::MyersCandidate const &member : launch.members)
            assignment[member.candidate] = 0u | ((uint32_t)launch.words << 8) | (member.lanes << 16) | (member.rows ? 0u : 0u);
    for (szb::HostSlot const &host : schedule.slots) {
        uint32_t const tier = host.family >= szb::FAM_S32_GLOBAL_LIN && host.family <= szb::FAM_S32_LOCAL_AFF ? 2u : 1u;
        uint32_t const code = tier | ((uint32_t)szb::kStrips[host.shape.strip_index] << 8) | ((uint32_t)host.shape.lanes << 16) |
                              ((uint32_t)host.family << 24);
        assignment[host.slot.c0] = code;
        if (host.slot.c1 != szb::kNoCandidate)
            assignment[host.slot.c1] = code, partner[host.slot.c0] = host.slot.c1, partner[host.slot.c1] = host.slot.c0;
    }
    for (size_t i = 0; i < schedule.order.size(); ++i) order[i] = schedule.order[i];
    *n_jobs = schedule.jobs.size();
    for (size_t j = 0; j < schedule.jobs.size() && j < jobs_capacity; ++j) {
        jobs[3 * j] = schedule.slots[schedule.jobs[j].slot].slot.c0;
        jobs[3 * j + 1] = schedule.jobs[j].q_begin, jobs[3 * j + 2] = schedule.jobs[j].q_end;
    }
    return sz_success_k;
}

SZS_B200_API sz_status_t szs_b200_device_scope_stream(szs_device_scope_t scope, void **stream, char const **error_message) {
    if (!scope || !stream) return szb::report(szb::fail(sz_status_unknown_k, "Invalid null pointer argument"), error_message);
    auto *typed = static_cast<Scope *>(scope);
    if (typed->kind == Scope::kCpu) return szb::report(szb::fail(sz_device_code_mismatch_k, "Device scope is CPU-only"), error_message);
    if (Failure failure = typed->ensure_stream()) return szb::report(failure, error_message);
    *stream = typed->stream;
    return szb::report(szb::ok(), error_message);
}

SZS_B200_API sz_status_t szs_b200_set_bit_parallel(void *engine, int enabled) {
    if (!engine) return sz_status_unknown_k;
    static_cast<Engine *>(engine)->bit_parallel = enabled != 0;
    return sz_success_k;
}

SZS_B200_API sz_status_t szs_b200_scores_u64tape_rows(int family, void *engine, szs_device_scope_t device,
                                                      sz_sequence_u64tape_t const *queries,
                                                      sz_sequence_u64tape_t const *candidates, sz_size_t row_begin,
                                                      sz_size_t row_end, void *results, sz_size_t results_row_stride,
                                                      char const **error_message) {
    if (!engine || !device || !queries)
        return szb::report(szb::fail(sz_status_unknown_k, "Engine, device scope and queries must not be null"), error_message);
    auto *typed = static_cast<Engine *>(engine);
    if ((int)typed->algorithm != family)
        return szb::report(szb::fail(sz_status_unknown_k, "Engine handle belongs to another algorithm"), error_message);
    InputView q {InputView::kTape64, queries}, c {InputView::kTape64, candidates};
    Failure failure = szb::cross(*typed, *static_cast<Scope *>(device), q, c, results, results_row_stride, row_begin,
                                 row_end, false);
    return szb::report(failure, error_message);
}

SZS_B200_API sz_status_t szs_b200_partition_rows(sz_u64_t const *lengths, sz_size_t count, sz_size_t parts, int symmetric,
                                                 sz_u64_t *bounds) {
    if ((!lengths && count) || !bounds || !parts) return sz_status_unknown_k;
    szb::partition_rows(lengths, count, parts, symmetric != 0, bounds);
    return sz_success_k;
}

SZS_B200_API sz_status_t szs_b200_device_group_init(sz_size_t const *gpu_devices, sz_size_t count,
                                                    szs_b200_device_group_t *group, char const **error_message) {
    if (!group) return szb::report(szb::fail(sz_status_unknown_k, "Group pointer is null"), error_message);
    size_t const visible = (size_t)szb::visible_device_count();
    if (!gpu_devices) count = visible; // every visible device
    if (!count) return szb::report(szb::fail(sz_missing_gpu_k, "No CUDA device is visible"), error_message);
    std::unique_ptr<szb::DeviceGroup> made(new (std::nothrow) szb::DeviceGroup());
    if (!made) return szb::report(szb::fail(sz_bad_alloc_k, "Failed to allocate the device group"), error_message);
    for (size_t i = 0; i < count; ++i) {
        size_t const device = gpu_devices ? gpu_devices[i] : i;
        if (device >= visible) return szb::report(szb::fail(sz_missing_gpu_k, "No such CUDA device"), error_message);
        std::unique_ptr<Scope> scope(new (std::nothrow) Scope());
        if (!scope) return szb::report(szb::fail(sz_bad_alloc_k, "Failed to allocate device scope"), error_message);
        scope->kind = Scope::kGpu, scope->cpu_cores = 0, scope->gpu_device = (int)device;
        made->scopes.push_back(std::move(scope));
    }
    try {
        made->start();
    }
    catch (...) {
        return szb::report(szb::fail(sz_bad_alloc_k, "Failed to start the device group's host threads"), error_message);
    }
    *group = made.release();
    return szb::report(szb::ok(), error_message);
}

SZS_B200_API sz_status_t szs_b200_device_group_size(szs_b200_device_group_t group, sz_size_t *count) {
    if (!group || !count) return sz_status_unknown_k;
    *count = static_cast<szb::DeviceGroup *>(group)->scopes.size();
    return sz_success_k;
}

SZS_B200_API void szs_b200_device_group_free(szs_b200_device_group_t group) {
    delete static_cast<szb::DeviceGroup *>(group);
}

SZS_B200_API sz_status_t szs_b200_scores_u64tape_sharded(int family, void *engine, szs_b200_device_group_t group,
                                                         sz_sequence_u64tape_t const *queries,
                                                         sz_sequence_u64tape_t const *candidates, void *results,
                                                         sz_size_t results_row_stride, char const **error_message) {
    if (!engine || !group || !queries)
        return szb::report(szb::fail(sz_status_unknown_k, "Engine, device group and queries must not be null"), error_message);
    auto *typed = static_cast<Engine *>(engine);
    if ((int)typed->algorithm != family)
        return szb::report(szb::fail(sz_status_unknown_k, "Engine handle belongs to another algorithm"), error_message);
    return szb::report(szb::sharded_cross(*typed, *static_cast<szb::DeviceGroup *>(group), queries, candidates, results,
                                          results_row_stride),
                       error_message);
}

SZS_B200_API sz_status_t szs_b200_comm_unique_id(void *unique_id, char const **error_message) {
    if (!unique_id) return szb::report(szb::fail(sz_status_unknown_k, "Identifier pointer is null"), error_message);
    szb::NcclApi const &nccl = szb::nccl_api();
    if (!nccl.ready()) return szb::report(szb::fail(sz_status_unknown_k, "libnccl.so.2 could not be loaded"), error_message);
    static_assert(sizeof(ncclUniqueId) == SZS_B200_COMM_ID_BYTES, "identifier size is part of the ABI");
    ncclResult_t const result = nccl.GetUniqueId(static_cast<ncclUniqueId *>(unique_id));
    if (result != ncclSuccess) return szb::report(szb::fail(sz_status_unknown_k, nccl.GetErrorString(result)), error_message);
    return szb::report(szb::ok(), error_message);
}

SZS_B200_API sz_status_t szs_b200_comm_init(void const *unique_id, int world_size, int rank, sz_size_t gpu_device,
                                            szs_b200_comm_t *comm, char const **error_message) {
    if (!unique_id || !comm || world_size < 1 || rank < 0 || rank >= world_size)
        return szb::report(szb::fail(sz_status_unknown_k, "Invalid communicator arguments"), error_message);
    szb::NcclApi const &nccl = szb::nccl_api();
    if (!nccl.ready()) return szb::report(szb::fail(sz_status_unknown_k, "libnccl.so.2 could not be loaded"), error_message);
    if ((size_t)szb::visible_device_count() <= gpu_device)
        return szb::report(szb::fail(sz_missing_gpu_k, "No such CUDA device"), error_message);
    std::unique_ptr<szb::Communicator> made(new (std::nothrow) szb::Communicator());
    if (!made) return szb::report(szb::fail(sz_bad_alloc_k, "Failed to allocate the communicator"), error_message);
    made->world = world_size, made->rank = rank, made->device = (int)gpu_device;
    if (cudaSetDevice((int)gpu_device) != cudaSuccess) {
        cudaGetLastError();
        return szb::report(szb::fail(sz_missing_gpu_k, "cudaSetDevice failed"), error_message);
    }
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    ncclResult_t const result = nccl.CommInitRank(&made->comm, world_size, id, rank);
    if (result != ncclSuccess) return szb::report(szb::fail(sz_status_unknown_k, nccl.GetErrorString(result)), error_message);
    *comm = made.release();
    return szb::report(szb::ok(), error_message);
}

SZS_B200_API void szs_b200_comm_free(szs_b200_comm_t comm) { delete static_cast<szb::Communicator *>(comm); }

SZS_B200_API sz_status_t szs_b200_comm_scores_u64tape(int family, void *engine, szs_device_scope_t device,
                                                      szs_b200_comm_t comm, sz_sequence_u64tape_t const *queries,
                                                      sz_sequence_u64tape_t const *candidates, void *results,
                                                      sz_size_t results_row_stride, char const **error_message) {
    if (!engine || !device || !comm)
        return szb::report(szb::fail(sz_status_unknown_k, "Engine, device scope and communicator must not be null"), error_message);
    auto *typed = static_cast<Engine *>(engine);
    if ((int)typed->algorithm != family)
        return szb::report(szb::fail(sz_status_unknown_k, "Engine handle belongs to another algorithm"), error_message);
    return szb::report(szb::communicator_cross(*typed, *static_cast<Scope *>(device), *static_cast<szb::Communicator *>(comm),
                                               queries, candidates, results, results_row_stride),
                       error_message);
}

} // extern "C"

#pragma endregion
