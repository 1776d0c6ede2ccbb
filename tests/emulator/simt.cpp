/**
 *  @file simt.cpp (tests/emulator)
 *  @brief TEST INFRASTRUCTURE: a lock-step SIMT emulator for the CPU test tier. One fiber per CUDA thread of the running
 *         CTA (hand-rolled x86-64 context switch, no system calls), CTAs of a grid one after the other, warp collectives
 *         and `__syncthreads` as rendezvous points with deadlock detection, and the handful of CUDA runtime entry points
 *         the host driver uses, over plain host memory.
 *
 *  Scheduling: the lanes of a warp run round-robin until each is finished or waits (for its warp or for the CTA); warps
 *  run in turn; when every unfinished thread of the CTA waits at the barrier, the barrier opens. A round in which nothing
 *  can move is a deadlock (a collective some lanes never reach) and aborts the process with a message.
 */
#include <cuda_runtime.h>

#include <sys/mman.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#if !defined(__x86_64__)
#error "the fiber switch below is written for x86-64"
#endif

/* `build.py --sanitize` (AddressSanitizer): device allocations become exact-size heap blocks, the part of the CTA's shared
 * memory a launch did not ask for is poisoned, and the fiber switches are announced to the sanitizer. */
#if defined(SIMT_SANITIZE)
#include <sanitizer/asan_interface.h>
#include <sanitizer/common_interface_defs.h>
#endif

extern "C" void simt_switch(void **save_stack_pointer, void *load_stack_pointer);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_switch,.-simt_switch
)");

namespace simt {

Index3 current_thread = {0, 0, 0}, current_block = {0, 0, 0}, current_block_dim = {1, 1, 1}, current_grid_dim = {1, 1, 1};

namespace {

constexpr size_t kStackBytes = 96 * 1024, kMaxThreads = 1024, kSharedBytes = 227 * 1024;

enum class Wait { kNone, kWarp, kCta, kWord, kDone };

struct Fiber {
    void *stack_pointer = nullptr;
    Wait wait = Wait::kNone;
    uint64_t generation = 0; // of the warp or CTA rendezvous this fiber waits for
    uint32_t const volatile *word = nullptr; // kWord: runnable once *word != value
    uint32_t value = 0;
};

struct Warp {
    uint32_t deposits[2][32];
    uint64_t tags[2][32]; // the rendezvous (generation) a deposit belongs to: lanes that left the kernel read as zero
    int kinds[2][32];
    unsigned arrived = 0, live = 0, phase = 0;
    uint64_t generation = 0;
};

struct Machine {
    char *stacks = nullptr;
    alignas(128) unsigned char shared[kSharedBytes];
    Fiber fibers[kMaxThreads];
    Warp warps[kMaxThreads / 32];
    uint32_t snapshots[kMaxThreads][32]; // what each thread saw in its latest warp rendezvous
    unsigned threads = 0, live_threads = 0, barrier_arrived = 0, running = 0;
    uint64_t barrier_generation = 0;
    void *scheduler_stack_pointer = nullptr;
    void const *scheduler_stack_bottom = nullptr; // sanitizer builds only
    size_t scheduler_stack_size = 0;
    void (*body)(void *) = nullptr;
    void *context = nullptr;
    bool inside_kernel = false;
};

Machine machine;
std::mutex launch_mutex; // host threads of a device group launch concurrently; the machine runs one grid at a time

[[noreturn]] void die(char const *what) {
    std::fprintf(stderr, "[simt emulator] %s (block %u, thread %u)\n", what, current_block.x, current_thread.x);
    std::abort();
}

void yield_to_scheduler() {
    Fiber &self = machine.fibers[machine.running];
#if defined(SIMT_SANITIZE)
    void *fake = nullptr;
    __sanitizer_start_switch_fiber(self.wait == Wait::kDone ? nullptr : &fake, machine.scheduler_stack_bottom, machine.scheduler_stack_size);
#endif
    simt_switch(&self.stack_pointer, machine.scheduler_stack_pointer);
#if defined(SIMT_SANITIZE)
    __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#endif
}

void leave_warp(Warp &warp);

void fiber_main() {
#if defined(SIMT_SANITIZE)
    __sanitizer_finish_switch_fiber(nullptr, &machine.scheduler_stack_bottom, &machine.scheduler_stack_size);
#endif
    machine.body(machine.context);
    Fiber &self = machine.fibers[machine.running];
    self.wait = Wait::kDone;
    machine.live_threads--;
    leave_warp(machine.warps[machine.running / 32]);
    // A thread that leaves the kernel counts as arrived at every later barrier (as on the hardware since Volta).
    if (machine.live_threads && machine.barrier_arrived == machine.live_threads)
        machine.barrier_arrived = 0, machine.barrier_generation++;
    yield_to_scheduler();
    die("a finished fiber was resumed");
}

void release_warp(Warp &warp) { warp.arrived = 0, warp.phase ^= 1, warp.generation++; }

void leave_warp(Warp &warp) {
    warp.live--;
    if (warp.live && warp.arrived == warp.live) release_warp(warp);
}

void prepare_fiber(unsigned index) {
    char *const top = machine.stacks + (size_t)(index + 1) * kStackBytes; // 16-byte aligned
    void **slots = reinterpret_cast<void **>(top) - 8;
    // Popped by simt_switch: r15, r14, r13, r12, rbx, rbp, then `ret` into fiber_main with rsp = top - 8 (as after a call).
    for (int i = 0; i < 6; ++i) slots[i] = nullptr;
    slots[6] = reinterpret_cast<void *>(&fiber_main);
    slots[7] = nullptr;
    machine.fibers[index].stack_pointer = slots;
    machine.fibers[index].wait = Wait::kNone;
}

bool runnable(unsigned index) {
    Fiber const &fiber = machine.fibers[index];
    switch (fiber.wait) {
    case Wait::kNone: return true;
    case Wait::kWarp: return machine.warps[index / 32].generation != fiber.generation;
    case Wait::kCta: return machine.barrier_generation != fiber.generation;
    case Wait::kWord: return *fiber.word != fiber.value;
    default: return false;
    }
}

void run_cta(unsigned block) {
    current_block = {block, 0, 0};
    unsigned const threads = machine.threads, n_warps = (threads + 31) / 32;
    machine.live_threads = threads, machine.barrier_arrived = 0;
    for (unsigned w = 0; w < n_warps; ++w) {
        Warp &warp = machine.warps[w];
        warp.arrived = 0, warp.phase = 0, warp.live = std::min(32u, threads - w * 32);
        for (int phase = 0; phase < 2; ++phase)
            for (int lane = 0; lane < 32; ++lane)
                warp.deposits[phase][lane] = 0, warp.kinds[phase][lane] = 0, warp.tags[phase][lane] = ~0ull;
    }
    for (unsigned t = 0; t < threads; ++t) prepare_fiber(t);
    // SZB_EMULATOR_SHUFFLE=<seed>: warps and lanes take their turns in a pseudo-random order that changes all the time.
    // Properly synchronised kernels compute the same thing under every order; one that leans on an accident of the default
    // order (a missing barrier, a read before a neighbour's write) shows up as a parity failure.
    static char const *const shuffle_text = getenv("SZB_EMULATOR_SHUFFLE");
    static uint64_t shuffle_state = shuffle_text ? strtoull(shuffle_text, nullptr, 10) * 2654435761ull + 1 : 0;
    auto next_random = [&]() -> unsigned {
        shuffle_state = shuffle_state * 6364136223846793005ull + 1442695040888963407ull;
        return (unsigned)(shuffle_state >> 33);
    };
    while (machine.live_threads) {
        bool moved = false;
        unsigned const warp_offset = shuffle_text ? next_random() % n_warps : 0, warp_stride = shuffle_text ? (next_random() | 1u) : 1;
        for (unsigned turn = 0; turn < n_warps; ++turn) {
            // (n_warps is a power of two for every kernel of the library, so an odd stride visits each warp once; other
            // sizes fall back to the plain order.)
            unsigned const w = (n_warps & (n_warps - 1)) ? turn : (warp_offset + turn * warp_stride) % n_warps;
            // Lanes of one warp take turns until none of them can move: most synchronisation is warp-level.
            for (bool warp_moved = true; warp_moved;) {
                warp_moved = false;
                unsigned const lane_offset = shuffle_text ? next_random() & 31 : 0, lane_stride = shuffle_text ? (next_random() | 1u) & 31 : 1;
                for (unsigned step = 0; step < 32; ++step) {
                    unsigned const t = w * 32 + ((lane_offset + step * lane_stride) & 31);
                    if (t >= threads) continue;
                    // With shuffling on, a runnable lane is sometimes passed over: neighbours then overtake it.
                    if (shuffle_text && (next_random() & 7) == 0) {
                        if (runnable(t)) warp_moved = moved = true;
                        continue;
                    }
                    if (!runnable(t)) continue;
                    machine.fibers[t].wait = Wait::kNone;
                    machine.running = t;
                    current_thread = {t, 0, 0};
#if defined(SIMT_SANITIZE)
                    void *fake = nullptr;
                    __sanitizer_start_switch_fiber(&fake, machine.stacks + (size_t)t * kStackBytes, kStackBytes);
#endif
                    simt_switch(&machine.scheduler_stack_pointer, machine.fibers[t].stack_pointer);
#if defined(SIMT_SANITIZE)
                    __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#endif
                    warp_moved = moved = true;
                }
                // ... and a warp sometimes yields to the others in the middle of its work (between two collectives).
                if (shuffle_text && (next_random() & 3) == 0) break;
            }
        }
        if (!moved && machine.live_threads) die("deadlock: no thread of the CTA can make progress");
    }
}

} // namespace

uint32_t *dynamic_shared() { return reinterpret_cast<uint32_t *>(machine.shared); }

void launch(unsigned grid, unsigned threads, size_t shared_bytes, void (*body)(void *), void *context) {
    std::lock_guard<std::mutex> lock(launch_mutex);
    if (machine.inside_kernel) die("nested kernel launch");
    if (!threads || threads > kMaxThreads) die("unsupported CTA size");
    if (shared_bytes > kSharedBytes) die("dynamic shared memory beyond 227 KB");
    if (!machine.stacks) {
        void *region = mmap(nullptr, kStackBytes * kMaxThreads, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (region == MAP_FAILED) die("cannot map the fiber stacks");
        machine.stacks = static_cast<char *>(region);
    }
    machine.threads = threads, machine.body = body, machine.context = context, machine.inside_kernel = true;
    current_block_dim = {threads, 1, 1}, current_grid_dim = {grid, 1, 1};
    static bool const trace = getenv("SZB_EMULATOR_TRACE") != nullptr;
    if (trace) std::fprintf(stderr, "[simt emulator] launch: %u CTAs of %u threads, %zu bytes of shared memory\n", grid, threads, shared_bytes);
#if defined(SIMT_SANITIZE)
    __asan_unpoison_memory_region(machine.shared, kSharedBytes);
    __asan_poison_memory_region(machine.shared + shared_bytes, kSharedBytes - shared_bytes);
#endif
    for (unsigned block = 0; block < grid; ++block) run_cta(block);
    if (trace) std::fprintf(stderr, "[simt emulator] launch finished\n");
    machine.inside_kernel = false;
}

void barrier_cta() {
    Fiber &self = machine.fibers[machine.running];
    if (++machine.barrier_arrived == machine.live_threads) {
        machine.barrier_arrived = 0, machine.barrier_generation++;
        return;
    }
    self.wait = Wait::kCta, self.generation = machine.barrier_generation;
    yield_to_scheduler();
}

void wait_while_equal(uint32_t const volatile *word, uint32_t value) {
    if (*word != value) return;
    Fiber &self = machine.fibers[machine.running];
    self.wait = Wait::kWord, self.word = word, self.value = value;
    yield_to_scheduler();
}

uint32_t const *exchange(uint32_t value, int kind) {
    unsigned const index = machine.running, lane = index & 31, first = index & ~31u;
    Warp &warp = machine.warps[index / 32];
    unsigned const phase = warp.phase;
    uint64_t const rendezvous = warp.generation;
    warp.deposits[phase][lane] = value, warp.kinds[phase][lane] = kind, warp.tags[phase][lane] = rendezvous;
    if (++warp.arrived == warp.live) {
        // Everybody is here: the collective must be the same call for all of them.
        for (unsigned other = 0; other < 32; ++other)
            if (warp.tags[phase][other] == rendezvous && warp.kinds[phase][other] != kind)
                die("lanes of one warp met in different collectives");
        release_warp(warp);
    }
    else {
        Fiber &self = machine.fibers[index];
        self.wait = Wait::kWarp, self.generation = rendezvous;
        yield_to_scheduler();
    }
    // Each thread keeps its own copy: a lane that has left the kernel, or runs ahead into the next rendezvous, must not
    // change what a slower lane is still about to read.
    uint32_t *const seen = machine.snapshots[index];
    for (unsigned other = 0; other < 32; ++other)
        seen[other] = warp.tags[phase][other] == rendezvous ? warp.deposits[phase][other] : 0;
    (void)first;
    return seen;
}

} // namespace simt

namespace szb {
void misaligned_shared_load(uint32_t address, uint32_t bytes) {
    std::fprintf(stderr, "[simt emulator] ld.shared of %u bytes at shared address %u is misaligned\n", bytes, address);
    std::abort();
}
void bulk_copy_alignment_violation() {
    std::fprintf(stderr, "[simt emulator] cp.async.bulk needs 16-byte aligned source, destination and size\n");
    std::abort();
}
} // namespace szb

/* ---- CUDA runtime over host memory ----------------------------------------------------------------------------------- */

namespace {

void api_trace(char const *name) {
    static bool const trace = getenv("SZB_EMULATOR_TRACE") != nullptr && atoi(getenv("SZB_EMULATOR_TRACE")) > 1;
    if (trace) std::fprintf(stderr, "[simt emulator] %s\n", name);
}

struct Allocation {
    size_t bytes;
    cudaMemoryType type;
};
std::map<uintptr_t, Allocation> allocations;
std::mutex allocations_mutex;

cudaError_t allocate(void **pointer, size_t bytes, cudaMemoryType type) {
    // Guard bytes on both sides: vector loads and bulk copies touch the 16-byte windows around a buffer's ends.
#if defined(SIMT_SANITIZE)
    size_t const rounded = bytes ? bytes : 1; // exact: the sanitizer's red zones sit right behind the last byte
    void *block = nullptr;
    if (posix_memalign(&block, 256, rounded) != 0) return cudaErrorMemoryAllocation;
    memset(block, 0xA5, rounded);
    *pointer = block;
#else
    size_t const rounded = (bytes + 255) & ~(size_t)255;
    void *block = nullptr;
    if (posix_memalign(&block, 256, rounded + 512) != 0) return cudaErrorMemoryAllocation;
    memset(block, 0xA5, rounded + 512); // reading uninitialised device memory should not look like zeros
    *pointer = static_cast<char *>(block) + 256;
#endif
    std::lock_guard<std::mutex> lock(allocations_mutex);
    allocations[reinterpret_cast<uintptr_t>(*pointer)] = {rounded, type};
    return cudaSuccess;
}

cudaError_t release(void *pointer) {
    if (!pointer) return cudaSuccess;
    {
        std::lock_guard<std::mutex> lock(allocations_mutex);
        if (!allocations.erase(reinterpret_cast<uintptr_t>(pointer))) return cudaErrorInvalidValue;
    }
#if defined(SIMT_SANITIZE)
    free(pointer);
#else
    free(static_cast<char *>(pointer) - 256);
#endif
    return cudaSuccess;
}

int emulated_sms() {
    char const *text = getenv("SZB_EMULATED_SMS");
    int const sms = text ? atoi(text) : 2;
    return sms > 0 ? sms : 2;
}

} // namespace

cudaError_t cudaGetLastError() { return cudaSuccess; }
char const *cudaGetErrorName(cudaError_t error) { return error == cudaSuccess ? "cudaSuccess" : "cudaErrorEmulated"; }
cudaError_t cudaGetDeviceCount(int *count) { return *count = 1, cudaSuccess; }
cudaError_t cudaSetDevice(int device) {
    api_trace("cudaSetDevice"); return device == 0 ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaGetDevice(int *device) { return *device = 0, cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *value, cudaDeviceAttr attribute, int device) {
    api_trace("cudaDeviceGetAttribute");
    if (device != 0) return cudaErrorInvalidDevice;
    switch (attribute) {
    case cudaDevAttrMultiProcessorCount: *value = emulated_sms(); break;
    case cudaDevAttrComputeCapabilityMajor: *value = 10; break;
    case cudaDevAttrComputeCapabilityMinor: *value = 0; break;
    default: return cudaErrorInvalidValue;
    }
    return cudaSuccess;
}
cudaError_t cudaMalloc(void **pointer, size_t bytes) {
    api_trace("cudaMalloc"); return allocate(pointer, bytes, cudaMemoryTypeDevice); }
cudaError_t cudaMallocHost(void **pointer, size_t bytes) {
    api_trace("cudaMallocHost"); return allocate(pointer, bytes, cudaMemoryTypeHost); }
cudaError_t cudaMallocManaged(void **pointer, size_t bytes) { return allocate(pointer, bytes, cudaMemoryTypeManaged); }
cudaError_t cudaFree(void *pointer) { return release(pointer); }
cudaError_t cudaFreeHost(void *pointer) { return release(pointer); }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *attributes, void const *pointer) {
    api_trace("cudaPointerGetAttributes");
    std::lock_guard<std::mutex> lock(allocations_mutex);
    uintptr_t const address = reinterpret_cast<uintptr_t>(pointer);
    auto after = allocations.upper_bound(address);
    if (after != allocations.begin()) {
        auto const &entry = *--after;
        if (address < entry.first + entry.second.bytes) {
            attributes->type = entry.second.type, attributes->device = 0;
            return cudaSuccess;
        }
    }
    attributes->type = cudaMemoryTypeUnregistered, attributes->device = -1;
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void *destination, void const *source, size_t bytes, cudaMemcpyKind) {
    api_trace("cudaMemcpy");
    if (bytes) memmove(destination, source, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *destination, void const *source, size_t bytes, cudaMemcpyKind kind, cudaStream_t) {
    return cudaMemcpy(destination, source, bytes, kind);
}
cudaError_t cudaMemcpy2DAsync(void *destination, size_t destination_pitch, void const *source, size_t source_pitch, size_t width,
                              size_t height, cudaMemcpyKind, cudaStream_t) {
    api_trace("cudaMemcpy2DAsync");
    for (size_t row = 0; row < height; ++row)
        memmove(static_cast<char *>(destination) + row * destination_pitch, static_cast<char const *>(source) + row * source_pitch, width);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *destination, int value, size_t bytes, cudaStream_t) {
    api_trace("cudaMemsetAsync");
    if (bytes) memset(destination, value, bytes);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *stream, unsigned) {
    api_trace("cudaStreamCreateWithFlags");
    return *stream = reinterpret_cast<cudaStream_t>(malloc(1)), cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t stream) { return free(stream), cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) {
    api_trace("cudaStreamSynchronize"); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *event) {
    api_trace("cudaEventCreate"); return *event = reinterpret_cast<cudaEvent_t>(malloc(1)), cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *event, unsigned) { return cudaEventCreate(event); }
cudaError_t cudaEventDestroy(cudaEvent_t event) { return free(event), cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) {
    api_trace("cudaEventRecord"); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *milliseconds, cudaEvent_t, cudaEvent_t) { return *milliseconds = 0.f, cudaSuccess; }

/* ---- Hooks for tests: "device" memory a Python test can fill and read back (what torch tensors do on a real GPU) ----- */

extern "C" __attribute__((visibility("default"))) void *simt_device_alloc(size_t bytes) {
    void *pointer = nullptr;
    return cudaMalloc(&pointer, bytes) == cudaSuccess ? pointer : nullptr;
}
extern "C" __attribute__((visibility("default"))) void simt_device_free(void *pointer) { cudaFree(pointer); }

