/**
 *  @file myers.cu
 *  @brief Instantiations of the bit-parallel unit-cost Levenshtein kernel, one per words-per-lane count.
 */
#include "myers.cuh"

namespace szb {

template <int W>
static cudaError_t launch_words(int grid, int threads, size_t smem, cudaStream_t stream, MyersArgs const &args) {
    auto kernel = myers_kernel<W>;
    static bool configured[64] = {false};
    int device = 0;
    cudaGetDevice(&device);
    if (device < 64 && !configured[device]) {
        cudaError_t status = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (status != cudaSuccess) return status;
        configured[device] = true;
    }
    kernel<<<grid, threads, smem, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t launch_myers(int words, int grid, int threads, size_t smem, cudaStream_t stream, MyersArgs const &args) {
    switch (words) {
    case 1: return launch_words<1>(grid, threads, smem, stream, args);
    case 2: return launch_words<2>(grid, threads, smem, stream, args);
    case 3: return launch_words<3>(grid, threads, smem, stream, args);
    case 4: return launch_words<4>(grid, threads, smem, stream, args);
    case 5: return launch_words<5>(grid, threads, smem, stream, args);
    case 6: return launch_words<6>(grid, threads, smem, stream, args);
    case 7: return launch_words<7>(grid, threads, smem, stream, args);
    case 8: return launch_words<8>(grid, threads, smem, stream, args);
    case 9: return launch_words<9>(grid, threads, smem, stream, args);
    default: return launch_words<10>(grid, threads, smem, stream, args);
    }
}

} // namespace szb
