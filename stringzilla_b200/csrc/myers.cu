/**
 *  @file myers.cu
 *  @brief Instantiations of the bit-parallel unit-cost Levenshtein kernels, one per words-per-lane count.
 */
#include "myers.cuh"

#include <atomic>

/** EXPERIMENT, off by default and not yet measured: every bit-parallel kernel asks for the same (largest) shared-memory
 *  carve-out, so that CTAs of launches with different words per lane may share an SM. */
#ifndef SZB_MYERS_ONE_CARVEOUT
#define SZB_MYERS_ONE_CARVEOUT 0
#endif

namespace szb {

template <int W, bool SHARED>
static cudaError_t launch_words(int grid, int threads, size_t smem, cudaStream_t stream, MyersArgs const &args) {
    auto kernel = SHARED ? myers_shared_kernel<W> : myers_kernel<W>;
    // Engines on different host threads launch concurrently: one atomic bit per device; devices past the mask are
    // configured on every launch.
    static std::atomic<uint64_t> configured {0};
    int device = 0;
    cudaGetDevice(&device);
    uint64_t const bit = device < 64 ? 1ull << device : 0;
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        cudaError_t status = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (status != cudaSuccess) return status;
        if (SZB_MYERS_ONE_CARVEOUT) {
            status = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (status != cudaSuccess) return status;
        }
        configured.fetch_or(bit, std::memory_order_release);
    }
    SZB_LAUNCH(kernel, grid, threads, smem, stream, args);
    return cudaGetLastError();
}

template <bool SHARED>
static cudaError_t launch_variant(int words, int grid, int threads, size_t smem, cudaStream_t stream,
                                  MyersArgs const &args) {
#ifdef SZB_ONLY_WORDS // kernel experiments: instantiate a single word count (tools/sass_mix.py)
    return launch_words<SZB_ONLY_WORDS, SHARED>(grid, threads, smem, stream, args);
#endif
    switch (words) {
    case 1: return launch_words<1, SHARED>(grid, threads, smem, stream, args);
    case 2: return launch_words<2, SHARED>(grid, threads, smem, stream, args);
    case 3: return launch_words<3, SHARED>(grid, threads, smem, stream, args);
    case 4: return launch_words<4, SHARED>(grid, threads, smem, stream, args);
    case 5: return launch_words<5, SHARED>(grid, threads, smem, stream, args);
    case 6: return launch_words<6, SHARED>(grid, threads, smem, stream, args);
    case 7: return launch_words<7, SHARED>(grid, threads, smem, stream, args);
    case 8: return launch_words<8, SHARED>(grid, threads, smem, stream, args);
    case 9: return launch_words<9, SHARED>(grid, threads, smem, stream, args);
    default: return launch_words<10, SHARED>(grid, threads, smem, stream, args);
    }
}

static cudaError_t launch_runes(int grid, int threads, size_t smem, cudaStream_t stream, MyersArgs const &args) {
    static std::atomic<uint64_t> configured {0};
    int device = 0;
    cudaGetDevice(&device);
    uint64_t const bit = device < 64 ? 1ull << device : 0;
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        cudaError_t status = cudaFuncSetAttribute(myers_rune_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (status != cudaSuccess) return status;
        configured.fetch_or(bit, std::memory_order_release);
    }
    SZB_LAUNCH(myers_rune_kernel<1>, grid, threads, smem, stream, args);
    return cudaGetLastError();
}

cudaError_t launch_myers(int words, int table, int grid, int threads, size_t smem, cudaStream_t stream, MyersArgs const &args) {
    if (table == kMyersRuneTable) return launch_runes(grid, threads, smem, stream, args);
    return table == kMyersSharedTable ? launch_variant<true>(words, grid, threads, smem, stream, args)
                                      : launch_variant<false>(words, grid, threads, smem, stream, args);
}

} // namespace szb
