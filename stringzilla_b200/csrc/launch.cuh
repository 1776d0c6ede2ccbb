/**
 *  @file launch.cuh
 *  @brief Host-visible launch table of the wavefront kernels. Each recurrence family is compiled in its own
 *         translation unit (`family.cu` with -DSZB_FAMILY=n) so the 40 instantiations build in parallel.
 */
#pragma once
#include "kernels.cuh"

namespace szb {

/** Strip shapes: (K columns per lane, G lanes per query). COLS = K * G DP columns per pass. */
struct StripShape {
    int k, g;
    constexpr int cols() const { return k * g; }
};
constexpr int kShapeCount = 5;
constexpr StripShape kShapes[kShapeCount] = {{8, 8}, {16, 8}, {32, 8}, {32, 16}, {32, 32}};

/** Picks the narrowest shape whose single pass covers `longest` columns; longer candidates take several passes. */
inline int shape_for_length(uint32_t longest) {
    for (int i = 0; i < kShapeCount; ++i)
        if (longest <= (uint32_t)kShapes[i].cols()) return i;
    return kShapeCount - 1;
}

/** Dynamic shared memory of one CTA: the profile, the 32x32 table and two control words. */
inline size_t shared_bytes(int shape, int n_classes, bool compare) {
    size_t const profile = compare ? 0 : (size_t)n_classes * kShapes[shape].cols() * 4;
    return profile + 1024 + 16;
}

inline bool family_is_affine(int family) { return family == 1 || family == 3 || family == 5 || family == 7; }

template <int FAM>
cudaError_t launch_family(int shape, int grid, size_t smem, cudaStream_t stream, KernelArgs const &args);

cudaError_t launch_wavefront(int family, int shape, int grid, size_t smem, cudaStream_t stream,
                             KernelArgs const &args);

} // namespace szb
