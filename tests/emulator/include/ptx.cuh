/**
 *  @file ptx.cuh (tests/emulator)
 *  @brief TEST INFRASTRUCTURE: the emulator's counterpart of `stringzilla_b200/csrc/ptx.cuh` — the same functions and
 *         macros, with C++ bodies that state what each PTX instruction computes. tests/emulator/build.py places THIS file
 *         next to symbolic links of the unchanged kernel sources, so that their `#include "ptx.cuh"` finds it.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#define SZB_LAUNCH(kernel, grid, threads, shared_bytes, stream, ...) \
    simt::launch((unsigned)(grid), (unsigned)(threads), (size_t)(shared_bytes), [&] { kernel(__VA_ARGS__); })

#define SZB_DYNAMIC_SHARED(name) uint32_t *const name = simt::dynamic_shared()

namespace szb {

/* mad.lo.u32 / mul.hi.u32 */
inline uint32_t add_on_fma_pipe(uint32_t a, uint32_t b, uint32_t one) { return a * one + b; }
inline uint32_t multiply_add_on_fma_pipe(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }
inline uint32_t multiply_high_on_fma_pipe(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t opaque_two() { return 2; }

/* The shared window: byte offsets from the start of the CTA's dynamic shared memory. */
inline uint32_t shared_window_address(void const *pointer) {
    return (uint32_t)((char const *)pointer - (char const *)simt::dynamic_shared());
}
void misaligned_shared_load(uint32_t address, uint32_t bytes);
inline uint32_t load_shared_word(uint32_t address) {
    if (address & 3u) misaligned_shared_load(address, 4); // the hardware faults on a misaligned ld.shared
    return *(uint32_t const *)((char const *)simt::dynamic_shared() + address);
}
inline void load_shared_pair(uint32_t address, uint32_t &first, uint32_t &second) {
    if (address & 7u) misaligned_shared_load(address, 8);
    first = load_shared_word(address), second = load_shared_word(address + 4);
}
inline void load_shared_quad(uint32_t address, uint32_t &first, uint32_t &second, uint32_t &third, uint32_t &fourth) {
    if (address & 15u) misaligned_shared_load(address, 16);
    first = load_shared_word(address), second = load_shared_word(address + 4);
    third = load_shared_word(address + 8), fourth = load_shared_word(address + 12);
}

/* cp.async.bulk + mbarrier. The 8 bytes of the barrier hold [0] the parity of the phase in progress and [1] the bytes still
 * in flight; the copy itself is immediate, and completes the phase once every expected byte has landed. Waiters block until
 * the parity moves on, so nobody reads the staged bytes before they are there. The copy checks what the hardware demands
 * of a bulk copy: 16-byte aligned source, destination and size. */
inline uint32_t *mbarrier_words(uint32_t barrier) { return (uint32_t *)((char *)simt::dynamic_shared() + barrier); }
inline void mbarrier_init(uint32_t barrier, uint32_t) { mbarrier_words(barrier)[0] = 0, mbarrier_words(barrier)[1] = 0; }
inline void mbarrier_init_fence() {}
inline void mbarrier_expect_bytes(uint32_t barrier, uint32_t bytes) { mbarrier_words(barrier)[1] += bytes; }
void bulk_copy_alignment_violation();
inline void bulk_copy_to_shared(uint32_t destination, void const *source, uint32_t bytes, uint32_t barrier) {
    if ((destination | bytes | (uint32_t)(uintptr_t)source) & 15u) bulk_copy_alignment_violation();
    memcpy((char *)simt::dynamic_shared() + destination, source, bytes);
    uint32_t *const words = mbarrier_words(barrier);
    if ((words[1] -= bytes) == 0) words[0] ^= 1u;
}
inline void mbarrier_wait(uint32_t barrier, uint32_t parity) { simt::wait_while_equal(&mbarrier_words(barrier)[0], parity); }

/* shfl.sync.up.b32 value|inside, ...: lanes without a left neighbour in their group receive `boundary`. */
template <uint32_t boundary>
inline uint32_t shuffle_up_or(uint32_t value, int group_lanes) {
    int const lane = (int)(threadIdx.x & 31);
    uint32_t const *const all = simt::exchange(value, 8);
    return (lane & (group_lanes - 1)) ? all[lane - 1] : boundary;
}

/* add.cc / addc.cc chain: bit 31 of `carry_in` enters the lowest word, the carry out of the top word is returned. */
template <int W>
inline uint32_t add_carry_chain(uint32_t (&sum)[W], uint32_t const (&t)[W], uint32_t const (&pv)[W], uint32_t carry_in) {
    uint64_t carry = carry_in >> 31;
    for (int w = 0; w < W; ++w) {
        uint64_t const wide = (uint64_t)t[w] + pv[w] + carry;
        sum[w] = (uint32_t)wide, carry = wide >> 32;
    }
    return (uint32_t)carry;
}

} // namespace szb
