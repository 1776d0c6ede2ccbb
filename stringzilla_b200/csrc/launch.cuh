/**
 *  @file launch.cuh
 *  @brief Host-visible launch table of the wavefront kernels. Each recurrence family is compiled in its own
 *         translation unit (`family.cu` with -DSZB_FAMILY=n) so the instantiations build in parallel.
 */
#pragma once
#include "kernels.cuh"

namespace szb {

/** Register-strip widths the kernels are instantiated for; a wavefront of G = 8, 16 or 32 lanes covers K * G columns. */
constexpr int kStripCount = 9;
constexpr int kStrips[kStripCount] = {8, 12, 16, 20, 24, 28, 32, 36, 40};
constexpr int kMaxCoverage = 40 * 32;

struct Shape {
    int strip_index; ///< index into kStrips
    int lanes;       ///< G
    int coverage() const { return kStrips[strip_index] * lanes; }
};

/**
 *  Picks the (K, G) whose coverage K * G fits `longest` columns most tightly: the cost of sweeping a query over a
 *  slot is proportional to the coverage, not to the candidate's length. Ties go to the wider strip (less per-step
 *  overhead). Candidates longer than the widest shape are split into equal passes. `min_lanes` = 32 serves calls too
 *  small to fill the device: the widest wavefront has the shortest dependent chain per step.
 */
inline Shape shape_for_length(uint32_t longest, int min_lanes = 8) {
    Shape best = {kStripCount - 1, 32};
    if (longest > (uint32_t)kMaxCoverage) {
        uint32_t const passes = (longest + kMaxCoverage - 1) / kMaxCoverage;
        uint32_t const width = (longest + passes - 1) / passes;
        for (int i = 0; i < kStripCount; ++i)
            if ((uint32_t)kStrips[i] * 32 >= width) return {i, 32};
        return best;
    }
    int best_coverage = kMaxCoverage + 1;
    for (int lanes = min_lanes; lanes <= 32; lanes *= 2)
        for (int i = 0; i < kStripCount; ++i) {
            int const coverage = kStrips[i] * lanes;
            if ((uint32_t)coverage < longest) continue;
            if (coverage < best_coverage || (coverage == best_coverage && kStrips[i] > kStrips[best.strip_index]))
                best = {i, lanes}, best_coverage = coverage;
        }
    return best;
}

/** Dynamic shared memory of one CTA: control words + mbarrier, the 32x32 table, the two TMA-staged candidate windows, and the profile: per class K / 4 vectors
 *  of 32 strip slots x 16 bytes (the slot count is fixed so that vector offsets are compile-time constants). */
inline size_t shared_bytes(int strip_index, int n_classes, bool compare) {
    size_t const profile = compare ? 0 : (size_t)n_classes * kStrips[strip_index] * 32 * 4;
    return 16 + 1024 + 2 * kStageBytes + profile;
}

constexpr bool family_is_affine(int family) { return family == 1 || family == 3 || family == 5 || family == 7; }

/** The widest register strip a family is instantiated for and the scheduler may pick. The 128-register budget sets it:
 *  packed affine recurrences keep two strips per lane (36 columns), the 32-bit tier adds the strip's symbols for compare
 *  mode (32 columns linear, 20 affine), and so does the packed byte-compare family (32). */
constexpr int widest_strip(int family) {
    if (family == FAM_LIN_COMPARE) return 32;
    if (family >= FAM_S32_GLOBAL_LIN) return family_is_affine(family) ? 20 : 32;
    return family_is_affine(family) ? 36 : 40;
}

template <int FAM>
cudaError_t launch_family(int strip_index, int grid, size_t smem, cudaStream_t stream, KernelArgs const &args);

cudaError_t launch_wavefront(int family, int strip_index, int grid, size_t smem, cudaStream_t stream,
                             KernelArgs const &args);

} // namespace szb
