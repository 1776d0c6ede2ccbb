/**
 *  @file myers.cuh
 *  @brief Bit-parallel unit-cost Levenshtein for sm_100a: Myers' / Hyyrö's vertical-delta recurrence, laid out as a
 *         skewed lane wavefront with a hardware carry chain.
 *
 *  The reference routes unit-cost byte-level Levenshtein to bit-parallel kernels on every backend
 *  (/root/reference/include/stringzillas/similarities/serial.hpp:2150-2300, cuda.cuh:2401-2852); its tier for
 *  |s| ~ 1024 gives one 64-bit word to each of 16 lanes and pays a 5-step Kogge-Stone shuffle scan for the carry of
 *  every text character (cuda.cuh:2767-2852). Here the carry never needs a scan: lane `l` of a group of G lanes owns W
 *  consecutive 32-bit words of the PATTERN (the candidate) and processes TEXT (query) character `s - l` at step `s`,
 *  so the three bits that cross a lane boundary (add carry, shifted-out Ph and Mh bits) were produced by the left
 *  neighbour one step earlier and arrive with a single `SHFL.UP`. Inside a lane the W words are chained by
 *  `ADDC.CC` (carry_chain.inc).
 *
 *  Per text character and 32-bit word (Hyyrö 2003, the formulation the reference restates at serial.hpp:2150-2211):
 *      Xv = Eq | Mv;  sum = (Eq & Pv) + Pv + carry;  Xh = (sum ^ Pv) | Eq;
 *      Ph = Mv | ~(Xh | Pv);  Mh = Pv & Xh;  Ph' = (Ph << 1) | ph_in;  Mh' = (Mh << 1) | mh_in;
 *      Pv = Mh' | ~(Xv | Ph');  Mv = Ph' & Xv
 *  = 7 LOP3 + 1 IADD3.X per 32 DP cells plus two 1-bit shifts, which alternate between SHF (ALU pipe) and
 *  IMAD / IMAD.HI (FMA pipe) from word to word. The distance is read off the final vertical deltas:
 *      D(m, n) = n + popc(Pv & mask_m) - popc(Mv & mask_m),
 *  so no per-step score tracking exists. `Eq` comes from a per-warp match table in shared memory,
 *  peq[class][word-in-lane][lane], built with `__ballot_sync` from the candidate's class bytes and replicated for the
 *  32 / G queries a warp sweeps at once, so every load is a conflict-free 128-byte wavefront.
 *
 *  carry_chain.inc is generated; the generator is the Python snippet kept in DESIGN.md §4.3.
 */
#pragma once
#include "kernels.cuh"

namespace szb {

#include "carry_chain.inc"

/** One candidate of a launch and the wavefront width chosen for it. */
struct MyersCandidate {
    uint32_t candidate, lanes;
};

struct MyersArgs {
    uint8_t const *q_cls;
    uint32_t const *q_off;
    uint8_t const *c_cls;
    uint32_t const *c_off;
    uint32_t const *q_order;
    MyersCandidate const *candidates; ///< longest first
    uint32_t n_candidates;
    uint32_t n_rows;                  ///< entries of q_order
    uint32_t blocks_per_candidate;    ///< task t = (candidate t / blocks, block t % blocks)
    uint32_t batches_per_block;       ///< a block is this many warp sweeps (each of 32 / lanes queries)
    uint32_t *task_counter;
    int32_t n_classes;
    int64_t *results;
    uint64_t stride;
    uint32_t row_base;
};

/**
 *  @brief Persistent warps: each pulls (candidate, query block) tasks from one atomic counter, builds the candidate's match table in its own
 *         slice of shared memory and sweeps the block's queries, 32 / G at a time.
 *  @tparam W 32-bit pattern words per lane; a group of G lanes covers 32 * W * G pattern characters.
 */
template <int W>
__global__ void __launch_bounds__(kThreads, 2) myers_kernel(MyersArgs const args) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr int ROW = 32 * W; // words per class row: [word-in-lane][lane]
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int const n_classes = args.n_classes;
    // Multipliers the compiler cannot see through, so that the shifts below stay integer multiply-adds.
    uint32_t two;
    asm volatile("mov.u32 %0, 2;" : "=r"(two));
    uint32_t *const peq = smem + (size_t)warp * n_classes * ROW;

    for (;;) {
        uint32_t task_index = 0;
        if (lane == 0) task_index = atomicAdd(args.task_counter, 1u);
        task_index = __shfl_sync(0xFFFFFFFFu, task_index, 0);
        if (task_index >= args.n_candidates * args.blocks_per_candidate) break;
        MyersCandidate const chosen = args.candidates[task_index / args.blocks_per_candidate];
        int const G = (int)chosen.lanes, GROUPS = 32 / G;
        uint32_t const block = task_index % args.blocks_per_candidate;
        uint32_t const q_first = block * args.batches_per_block * GROUPS;
        if (q_first >= args.n_rows) continue; // wider wavefronts need more blocks than narrower ones
        uint32_t const q_last = min(args.n_rows, q_first + args.batches_per_block * GROUPS);
        uint32_t const candidate = chosen.candidate;
        uint32_t const c_begin = args.c_off[candidate];
        int const m = (int)(args.c_off[candidate + 1] - c_begin); // pattern length
        uint8_t const *const c_ptr = args.c_cls + c_begin;
        int const gl = lane & (G - 1), gi = lane / G;

        // 1. Match table of the candidate: bit i of word p is set in row `a` iff candidate[32 p + i] has class a.
        __syncwarp();
        for (int p = 0; p < G * W; ++p) {
            int const position = p * 32 + lane;
            uint32_t const symbol = position < m ? c_ptr[position] : 0xFFFFFFFFu;
            int const word_in_lane = p % W, owner = p / W; // the lane (within a group) that owns pattern word p
            for (int a = 0; a < n_classes; ++a) {
                uint32_t const bits = __ballot_sync(0xFFFFFFFFu, symbol == (uint32_t)a);
                if (lane < GROUPS) peq[a * ROW + word_in_lane * 32 + lane * G + owner] = bits;
            }
        }
        __syncwarp();

        // Valid-bit masks of this lane's pattern words, for the final popcounts.
        uint32_t mask[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
            int const valid_bits = min(max(m - 32 * (gl * W + w), 0), 32);
            mask[w] = valid_bits == 32 ? 0xFFFFFFFFu : ((1u << valid_bits) - 1u);
        }

        // 2. Sweep the task's queries.
        for (uint32_t base = q_first; base < q_last; base += GROUPS) {
            uint32_t const q_pos = base + gi;
            bool const q_valid = q_pos < q_last;
            uint32_t const q = q_valid ? args.q_order[q_pos] : 0;
            uint32_t const q_begin_off = q_valid ? args.q_off[q] : 0;
            int const n = q_valid ? (int)(args.q_off[q + 1] - q_begin_off) : 0; // text length
            uint8_t const *const q_ptr = args.q_cls + q_begin_off;

            uint32_t pv[W], mv[W];
#pragma unroll
            for (int w = 0; w < W; ++w) pv[w] = 0xFFFFFFFFu, mv[w] = 0;

            int steps = n > 0 ? n + G - 1 : 0;
            for (int offset = 16; offset >= G; offset >>= 1)
                steps = max(steps, __shfl_xor_sync(0xFFFFFFFFu, steps, offset));

            // Bits crossing the lane boundary, packed: bit 31 = Ph carry, bit 30 = Mh carry, bit 0 = add carry.
            uint32_t out = 0;
            uint32_t a_next = n > 0 ? q_ptr[0] : 0; // only lane 0 of a group consumes it at step 0
            bool valid_next = gl == 0 && n > 0;
            for (int s = 0; s < steps; ++s) {
                uint32_t in = __shfl_up_sync(0xFFFFFFFFu, out, 1, G);
                if (gl == 0) in = 0x80000000u; // row 0 of the DP: D(0, j) - D(0, j - 1) = +1, i.e. a Ph carry
                bool const valid = valid_next;
                uint32_t const a = a_next;
                {   // Prefetch the class of the next text character of this lane.
                    int const next = s + 1 - gl;
                    valid_next = (unsigned)next < (unsigned)n;
                    if (valid_next) a_next = q_ptr[next];
                }
                if (valid) {
                    uint32_t const *const row = peq + a * ROW + lane;
                    uint32_t eq[W], t[W], sum[W];
#pragma unroll
                    for (int w = 0; w < W; ++w) eq[w] = row[w * 32], t[w] = eq[w] & pv[w];
                    uint32_t const carry = add_carry_chain<W>(sum, t, pv, in & 1u);
                    // The 1-bit left shifts of Ph and Mh alternate between the two integer pipes: even words use a
                    // funnel shift (ALU pipe), odd words `x * 2 + carry` with the carry from `hi(x * 2)` (IMAD and
                    // IMAD.HI on the FMA pipe), which keeps both pipes busy without saturating either.
                    uint32_t ph_below = in, mh_below = in << 1; // only bit 31 of each is consumed
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        uint32_t const xh = (sum[w] ^ pv[w]) | eq[w];
                        uint32_t const ph = mv[w] | ~(xh | pv[w]);
                        uint32_t const mh = pv[w] & xh;
                        uint32_t const xv = eq[w] | mv[w];
                        uint32_t ph_shifted, mh_shifted;
                        if (w & 1) {
                            uint32_t ph_carry, mh_carry;
                            asm("mul.hi.u32 %0, %1, %2;" : "=r"(ph_carry) : "r"(ph_below), "r"(two));
                            asm("mul.hi.u32 %0, %1, %2;" : "=r"(mh_carry) : "r"(mh_below), "r"(two));
                            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(ph_shifted) : "r"(ph), "r"(two), "r"(ph_carry));
                            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(mh_shifted) : "r"(mh), "r"(two), "r"(mh_carry));
                        }
                        else {
                            ph_shifted = __funnelshift_l(ph_below, ph, 1);
                            mh_shifted = __funnelshift_l(mh_below, mh, 1);
                        }
                        ph_below = ph, mh_below = mh;
                        pv[w] = mh_shifted | ~(xv | ph_shifted);
                        mv[w] = ph_shifted & xv;
                    }
                    out = (ph_below & 0x80000000u) | ((mh_below >> 1) & 0x40000000u) | carry;
                }
            }

            // 3. D(m, n) = n + sum over pattern bits of (Pv - Mv), reduced over the group's lanes.
            int delta = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) delta += __popc(pv[w] & mask[w]) - __popc(mv[w] & mask[w]);
            for (int offset = G / 2; offset >= 1; offset >>= 1) delta += __shfl_xor_sync(0xFFFFFFFFu, delta, offset);
            if (gl == 0 && q_valid)
                args.results[(uint64_t)(q - args.row_base) * args.stride + candidate] = (int64_t)n + delta;
        }
    }
}

constexpr int kMyersMaxWords = 10; // two 512-thread CTAs per SM need <= 64 registers per thread

/** Picks (W <= max_words, G) with the tightest coverage 32 * W * G >= pattern length; W = 0 when nothing fits. */
inline void myers_shape_for_length(uint32_t length, int max_words, int &words, int &lanes) {
    words = 0, lanes = 0;
    uint32_t best = 0xFFFFFFFFu;
    for (int g = 1; g <= 32; g *= 2)
        for (int w = 1; w <= kMyersMaxWords && w <= max_words; ++w) {
            uint32_t const coverage = 32u * w * g;
            if (coverage < length) continue;
            // Ties go to more words per lane: fewer lanes per query means more queries per warp and less overhead.
            if (coverage < best || (coverage == best && w > words)) best = coverage, words = w, lanes = g;
        }
}

cudaError_t launch_myers(int words, int grid, int threads, size_t smem, cudaStream_t stream, MyersArgs const &args);

} // namespace szb
