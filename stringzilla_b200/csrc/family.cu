/**
 *  @file family.cu
 *  @brief One translation unit per recurrence family: compile with -DSZB_FAMILY=<0..7>.
 */
#include "launch.cuh"

#include <atomic>

#ifndef SZB_FAMILY
#error "compile with -DSZB_FAMILY=<family index>"
#endif

namespace szb {

template <int FAM, int STRIP>
static cudaError_t launch_reachable_strip(int grid, size_t smem, cudaStream_t stream, KernelArgs const &args);

template <int FAM, int STRIP>
static cudaError_t launch_strip(int grid, size_t smem, cudaStream_t stream, KernelArgs const &args) {
    // Strips beyond the family's register budget are never scheduled (widest_strip, launch.cuh) and not instantiated.
    if constexpr (kStrips[STRIP] > widest_strip(FAM)) return cudaErrorInvalidValue;
    else return launch_reachable_strip<FAM, STRIP>(grid, smem, stream, args);
}

template <int FAM, int STRIP>
static cudaError_t launch_reachable_strip(int grid, size_t smem, cudaStream_t stream, KernelArgs const &args) {
    auto kernel = wavefront_kernel<FAM, kStrips[STRIP]>;
    // Opt in to the large dynamic shared-memory carve-out once per kernel and device.
    // Engines on different host threads launch concurrently: one atomic bit per device; devices past the mask are
    // configured on every launch.
    static std::atomic<uint64_t> configured {0};
    int device = 0;
    cudaGetDevice(&device);
    uint64_t const bit = device < 64 ? 1ull << device : 0;
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        cudaError_t status = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (status != cudaSuccess) return status;
        configured.fetch_or(bit, std::memory_order_release);
    }
    SZB_LAUNCH(kernel, grid, kThreads, smem, stream, args);
    return cudaGetLastError();
}

template <>
cudaError_t launch_family<SZB_FAMILY>(int strip, int grid, size_t smem, cudaStream_t stream, KernelArgs const &args) {
#ifdef SZB_ONLY_STRIP // kernel experiments: instantiate a single strip width (tools/sass_mix.py)
    return launch_strip<SZB_FAMILY, SZB_ONLY_STRIP>(grid, smem, stream, args);
#endif
    switch (strip) {
    case 0: return launch_strip<SZB_FAMILY, 0>(grid, smem, stream, args);
    case 1: return launch_strip<SZB_FAMILY, 1>(grid, smem, stream, args);
    case 2: return launch_strip<SZB_FAMILY, 2>(grid, smem, stream, args);
    case 3: return launch_strip<SZB_FAMILY, 3>(grid, smem, stream, args);
    case 4: return launch_strip<SZB_FAMILY, 4>(grid, smem, stream, args);
    case 5: return launch_strip<SZB_FAMILY, 5>(grid, smem, stream, args);
    case 6: return launch_strip<SZB_FAMILY, 6>(grid, smem, stream, args);
    case 7: return launch_strip<SZB_FAMILY, 7>(grid, smem, stream, args);
    default: return launch_strip<SZB_FAMILY, 8>(grid, smem, stream, args);
    }
}

} // namespace szb
