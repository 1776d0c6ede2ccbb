/**
 *  @file ptx.cuh
 *  @brief Every piece of inline PTX the kernels use, and the kernel-launch syntax, behind plain functions and one macro:
 *         pipe-pinned integer arithmetic, shared-window loads at byte addresses, the bulk-copy (TMA) engine with its
 *         mbarrier, a shuffle whose out-of-range predicate supplies a boundary value, and the hardware carry chain.
 *
 *  Keeping them in one header is what lets `tests/emulator/` compile the UNCHANGED kernels and host driver against a
 *  lock-step SIMT emulator on the CPU test tier (its own `ptx.cuh` and `cuda_runtime.h` come first on the include path
 *  there). Nothing in this file is reachable from that build, and nothing of the emulator is part of the library.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

/** `kernel<<<grid, threads, shared_bytes, stream>>>(arguments...)`. */
#define SZB_LAUNCH(kernel, grid, threads, shared_bytes, stream, ...) kernel<<<(grid), (threads), (shared_bytes), (stream)>>>(__VA_ARGS__)

/** The CTA's dynamic shared memory as an array of 32-bit words called `name`. */
#define SZB_DYNAMIC_SHARED(name) extern __shared__ __align__(16) uint32_t name[]

namespace szb {

/* ---- Integer arithmetic pinned to the FMA pipe ------------------------------------------------------------------- */

/** A 32-bit add that is guaranteed to issue on the FMA pipe (IMAD with a multiplier the compiler cannot fold): the
 *  packed recurrences saturate the ALU pipe with DPX instructions, and ptxas otherwise turns part of the plain adds
 *  into ALU-pipe VIADDs. */
__device__ __forceinline__ uint32_t add_on_fma_pipe(uint32_t a, uint32_t b, uint32_t one) {
    uint32_t sum;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(sum) : "r"(a), "r"(one), "r"(b));
    return sum;
}

/** `a * b + c` and the high word of `a * b` as IMAD / IMAD.HI: with b == 2 a 1-bit left shift with carry-in, and the bit
 *  shifted out. */
__device__ __forceinline__ uint32_t multiply_add_on_fma_pipe(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t result;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(result) : "r"(a), "r"(b), "r"(c));
    return result;
}
__device__ __forceinline__ uint32_t multiply_high_on_fma_pipe(uint32_t a, uint32_t b) {
    uint32_t result;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(result) : "r"(a), "r"(b));
    return result;
}

/** The constant 2 as a value the compiler cannot see through, so that `x * two` stays an integer multiply-add. */
__device__ __forceinline__ uint32_t opaque_two() {
    uint32_t two;
    asm volatile("mov.u32 %0, 2;" : "=r"(two));
    return two;
}

/* ---- Shared memory by byte address of the shared window ----------------------------------------------------------- */

__device__ __forceinline__ uint32_t shared_window_address(void const *pointer) { return (uint32_t)__cvta_generic_to_shared(pointer); }

__device__ __forceinline__ uint32_t load_shared_word(uint32_t address) {
    uint32_t word;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(address));
    return word;
}
__device__ __forceinline__ void load_shared_pair(uint32_t address, uint32_t &first, uint32_t &second) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(first), "=r"(second) : "r"(address));
}
__device__ __forceinline__ void load_shared_quad(uint32_t address, uint32_t &first, uint32_t &second, uint32_t &third, uint32_t &fourth) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(first), "=r"(second), "=r"(third), "=r"(fourth) : "r"(address));
}

/* ---- TMA (bulk asynchronous copy) into shared memory, completion on an mbarrier ------------------------------------ */

__device__ __forceinline__ void mbarrier_init(uint32_t barrier, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(barrier), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbarrier_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbarrier_expect_bytes(uint32_t barrier, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barrier), "r"(bytes) : "memory");
}
/** One 1-D bulk copy global -> shared (`cp.async.bulk`, the TMA engine without a tensor map: these are byte runs).
 *  Source, destination and size are multiples of 16 bytes. */
__device__ __forceinline__ void bulk_copy_to_shared(uint32_t destination, void const *source, uint32_t bytes, uint32_t barrier) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(destination),
                 "l"(source), "r"(bytes), "r"(barrier)
                 : "memory");
}
__device__ __forceinline__ void mbarrier_wait(uint32_t barrier, uint32_t parity) {
    asm volatile("{\n\t.reg .pred arrived;\n\t"
                 "SZB_WAIT_FOR_BYTES:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 arrived, [%0], %1;\n\t"
                 "@!arrived bra SZB_WAIT_FOR_BYTES;\n\t}" ::"r"(barrier),
                 "r"(parity)
                 : "memory");
}

/* ---- Shuffles ------------------------------------------------------------------------------------------------------ */

/** `__shfl_up_sync(full mask, value, 1, group_lanes)` whose out-of-range predicate supplies `boundary` to the first lane of
 *  every group (instead of handing it its own value back). */
template <uint32_t boundary>
__device__ __forceinline__ uint32_t shuffle_up_or(uint32_t value, int group_lanes) {
    uint32_t received;
    asm volatile("{\n\t.reg .pred inside;\n\t"
                 "shfl.sync.up.b32 %0|inside, %1, 1, %2, 0xffffffff;\n\t"
                 "@!inside mov.b32 %0, %3;\n\t}"
                 : "=r"(received)
                 : "r"(value), "r"((uint32_t)((32 - group_lanes) << 8)), "n"(boundary));
    return received;
}

/* ---- sum[w] = t[w] + pv[w] + carry over W words on the hardware carry flag ------------------------------------------- */

#include "carry_chain.inc"

} // namespace szb
