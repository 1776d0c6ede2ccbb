/**
 *  @file kernels.cuh
 *  @brief sm_100a device code of the batched string-similarity hot path: one skewed register wavefront per
 *         (query, candidate-pair), DPX min/max-fused recurrences on packed 2 x u16 cells.
 *
 *  Replaces (does not port) the reference's anti-diagonal shared-memory warp kernels and their Hopper DPX cell
 *  functors: /root/reference/include/stringzillas/similarities/cuda.cuh:1246-1626 and hopper.cuh:45-617.
 *
 *  Data flow of one launch
 *  -----------------------
 *  A "slot" is a pair of candidates (c0, c1) of similar length whose DP cells travel together as the low / high
 *  16-bit halves of one 32-bit register, so that every DPX instruction (VIMNMX3.U16x2 / VIADDMNMX.U16x2) retires two
 *  independent cells. A "job" is a slot x a block of queries. A persistent CTA pulls jobs from an atomic counter;
 *  for each job (and each pass of K * G columns, for candidates longer than the widest shape) it
 *    1. builds the slot's *substitution profile* in shared memory: prof[a][j] = packed (e(a, c0[j]), e(a, c1[j]))
 *       for every row class `a` — so the inner loop never compares characters or touches the 32x32 cost table,
 *    2. lets each warp (or sub-warp group of G lanes) sweep one query over the profile: lane `l` owns K consecutive
 *       DP columns in registers and processes DP row `s - l` at step `s` (a skewed wavefront); the only inter-lane
 *       traffic is one `__shfl_up` of the strip's last cell per step (two for affine gaps).
 *  K is a template parameter (the strip lives in registers), G is a run-time property of the job: the host picks
 *  the (K, G) whose coverage K * G fits the slot's length most tightly, because the cost of a sweep is proportional
 *  to K * G, not to the candidate's length.
 *  Profile rows are laid out [vector v][32 strip slots][16 bytes]: lane `l` reads its v-th vector at the compile-time
 *  offset v * 512, and the LDS.128 of the 8 lanes of a quarter-warp covers all 32 banks whatever rows (classes) the
 *  lanes are reading: conflict-free by construction.
 *
 *  Recurrences (all MAXIMISING; Levenshtein runs with negated costs and a negated result)
 *  --------------------------------------------------------------------------------------
 *  FAM_LIN_SHIFT  (Levenshtein & Needleman-Wunsch, linear gaps): u(i,j) = H(i,j) - (i+j)*g turns the gap term into
 *      a plain max:  u = max3(u_diag + e, u_up, u_left),  e = max(S - 2g, 0)  (S < 2g can never win against two
 *      gaps, so the clamp is exact). One IADD (FMA-pipe IMAD.IADD) + one VIMNMX3.U16x2 per packed cell pair, all
 *      boundaries are zero. Follows the semantics of reference test/similarities.cuh:47-100.
 *  FAM_AFF_SHIFT  (Levenshtein & NW, affine gaps): shifted by the extension cost and offset by O = 4*go,
 *      go = extend - open >= 0:  F = max(w_up - go, F_up), E = max(w_left - go, E_left), w = max3(w_diag + e, E, F),
 *      e = max(S, 2*open) - 2*extend.  Reference test/similarities.cuh:136-228 (Gotoh with the finite
 *      "discard" seeds, equivalent to -inf seeds whenever extend <= 0, which the host checks).
 *  FAM_SW_LIN / FAM_SW_AFF (Smith-Waterman): unshifted, cells biased by B so every half stays non-negative and a
 *      plain 32-bit add of an arithmetically packed (possibly negative) profile entry is exact per half.
 *      Reference test/similarities.cuh:103-133, 234-280.
 *  FAM_LIN_COMPARE  FAM_LIN_SHIFT for raw-byte alphabets of more than 31 symbols: the substitution term comes from a
 *      packed XOR of the strip's bytes (kept in registers) with the row byte instead of a profile read.
 *  FAM_S32_*      one candidate per register, signed 32-bit cells, literal baseline recurrences with DPX s32 ops;
 *      serves pairs whose scores do not fit 16 bits, unusual cost signs, affine-gap Levenshtein over large alphabets
 *      and UTF-32 symbol streams (compare mode: the substitution is a symbol comparison instead of a profile lookup).
 */
#pragma once
#include "ptx.cuh"

/** Smith-Waterman affine, variant: cells stored as H (not H + open), the `+ open` of both gap tracks formed by adds off
 *  the dependency chain (4 FMA-pipe adds per cell pair instead of 3, a 2-op instead of a 3-op chain per cell). */
#ifndef SZB_SW_SHORT_CHAIN
#define SZB_SW_SHORT_CHAIN 0
#endif

/** Two sub-strips per lane in the affine packed kernels (see `sweep_batch`): measured slower, kept as an experiment. */
#ifndef SZB_SPLIT_STRIPS
#define SZB_SPLIT_STRIPS 0
#endif

namespace szb {

enum Family : int {
    FAM_LIN_SHIFT = 0,
    FAM_AFF_SHIFT = 1,
    FAM_SW_LIN = 2,
    FAM_SW_AFF = 3,
    FAM_S32_GLOBAL_LIN = 4,
    FAM_S32_GLOBAL_AFF = 5,
    FAM_S32_LOCAL_LIN = 6,
    FAM_S32_LOCAL_AFF = 7,
    FAM_LIN_COMPARE = 8, ///< FAM_LIN_SHIFT whose substitution term is a packed byte comparison instead of a profile read
};
constexpr int kFamilyCount = 9;

constexpr int kThreads = 512;                  ///< 16 warps per persistent CTA, one CTA per SM
constexpr uint32_t kNoCandidate = 0xFFFFFFFFu; ///< `CandSlot::c1` of an unpaired slot

/** Scoring model of one engine, in the kernel's maximising convention. Lives in device memory. */
struct Scoring {
    int8_t sub[32 * 32]; ///< sub[a * 32 + b]: class(query byte) = a (row), class(candidate byte) = b (column)
    int32_t n_classes;   ///< profile rows actually in use (max class id + 1)
    int32_t open;        ///< cost ADDED for the first character of a gap (linear gaps: open == extend)
    int32_t extend;      ///< cost ADDED for every further character of a gap
    int32_t negate;      ///< 1 for Levenshtein: reported distance = -H
    int32_t bias;        ///< FAM_AFF_SHIFT: O = 4*go; FAM_SW_*: B; else 0
    int32_t compare;     ///< FAM_S32_*: 1 = strings hold raw bytes, substitution = (a == b) ? match : mismatch
    int32_t match;       ///< compare mode, maximising convention
    int32_t mismatch;    ///< compare mode, maximising convention
};

/** A pair of candidates travelling in the two 16-bit halves of each register (or a single one for FAM_S32_*). */
struct CandSlot {
    uint32_t c0, c1;     ///< candidate indices; c1 == kNoCandidate when unpaired
    uint32_t len0, len1; ///< lengths in bytes; len1 == 0 when unpaired
};

/** A slot scored against the queries `q_order[q_begin .. q_end)` by wavefronts of `lanes` lanes per query. */
struct Job {
    uint32_t slot, q_begin, q_end;
    uint32_t lanes; ///< G: 8, 16 or 32 lanes sweep one query; a warp sweeps 32 / G queries at a time
};

struct KernelArgs {
    uint8_t const *q_cls;  ///< class-mapped (or raw, in compare mode) query bytes
    uint32_t const *q_off; ///< Q+1 offsets into q_cls
    uint8_t const *c_cls;
    uint32_t const *c_off;
    uint32_t const *q_order; ///< query indices, longest first
    CandSlot const *slots;
    Job const *jobs;
    uint32_t n_jobs;
    uint32_t *job_counter;
    Scoring const *scoring;
    int64_t *results;        ///< results[(q - row_base) * stride + c]
    uint64_t stride;
    uint32_t row_base;
    uint32_t *scratch;       ///< per-CTA multi-pass boundary storage
    uint64_t scratch_words_per_cta;
    uint32_t wide_symbols;   ///< compare mode only: q_cls / c_cls are uint32 symbol streams (UTF-8 runes), offsets in symbols
};

__device__ __forceinline__ uint32_t pack_arith(int lo, int hi) { return (uint32_t)(hi * 65536 + lo); }
__device__ __forceinline__ uint32_t pack_bits(int v) { return ((uint32_t)v & 0xFFFFu) * 0x10001u; }

template <int FAM>
struct family_traits {
    static constexpr bool packed = FAM <= FAM_SW_AFF || FAM == FAM_LIN_COMPARE;
    static constexpr bool linear_shift = FAM == FAM_LIN_SHIFT || FAM == FAM_LIN_COMPARE;
    static constexpr bool affine =
        FAM == FAM_AFF_SHIFT || FAM == FAM_SW_AFF || FAM == FAM_S32_GLOBAL_AFF || FAM == FAM_S32_LOCAL_AFF;
    static constexpr bool local =
        FAM == FAM_SW_LIN || FAM == FAM_SW_AFF || FAM == FAM_S32_LOCAL_LIN || FAM == FAM_S32_LOCAL_AFF;
};

/** Substitution term stored in the profile for row class `a` against candidate class `b` (or padding). */
template <int FAM>
__device__ __forceinline__ int profile_term(int8_t const *sub, int a, int b, bool valid, int open, int extend,
                                            int bias) {
    int const s = valid ? (int)sub[a * 32 + b] : 0;
    if constexpr (family_traits<FAM>::linear_shift) { return valid ? max(s - 2 * open, 0) : 0; }
    else if constexpr (FAM == FAM_AFF_SHIFT) { return valid ? max(s, 2 * open) - 2 * extend : 2 * (open - extend); }
    else if constexpr (FAM == FAM_SW_LIN) { return valid ? s : -bias; }
    else if constexpr (FAM == FAM_SW_AFF) { // cells are kept as H + open (unless SZB_SW_SHORT_CHAIN)
        return SZB_SW_SHORT_CHAIN ? (valid ? s : -bias) : (valid ? s - open : -open - bias);
    }
    else { return valid ? s : -(1 << 28); }
}

/** Bytes reserved per staged candidate window (the slot's candidate bytes reach shared memory through the TMA engine,
 *  `bulk_copy_to_shared` in ptx.cuh): the widest pass (40 x 32 columns) plus the 16-byte alignment slack at both ends. */
constexpr int kStageBytes = 40 * 32 + 32;

/** Score of aligning `len` characters against nothing: serial.hpp:165-177 of the reference. */
template <int FAM>
__device__ __forceinline__ int64_t degenerate_score(int64_t len, int open, int extend) {
    if constexpr (family_traits<FAM>::local) return 0;
    else if constexpr (family_traits<FAM>::affine) return len == 0 ? 0 : (int64_t)open + (int64_t)extend * (len - 1);
    else return (int64_t)open * len;
}

/** Per-CTA constants of the recurrences, packed per half where a DPX instruction needs them bitwise. */
struct SweepConstants {
    int open, extend, bias, go, negate;
    int cmp_match, cmp_mismatch;
    bool compare;
    uint32_t neg_go_bits;  // FAM_AFF_SHIFT: "- go" for VIADDMNMX
    uint32_t bias_bits;    // FAM_SW_*: the zero floor in biased units
    uint32_t gap_bits;     // FAM_SW_*: "+ open" for VIADDMNMX
    uint32_t gap_arith;    // FAM_SW_LIN: "+ g" through a plain add
    uint32_t ext_arith;    // FAM_SW_AFF: "+ extend" through a plain add
    uint32_t cmp_match_bits; // FAM_LIN_SHIFT compare mode: e(match) in both halves
    uint32_t cmp_drop_bits;  // FAM_LIN_SHIFT compare mode: e(match) - e(mismatch) >= 0 in both halves
    uint32_t ext_bits;       // FAM_SW_AFF: "+ extend" for VIADDMNMX
    uint32_t open_arith;     // FAM_SW_AFF: "+ open" through a plain add
    uint32_t one;            // the opaque multiplier of add_on_fma_pipe
};

/**
 *  @brief Sweeps one batch of 32 / G queries over the profile of one pass of a slot.
 *  @tparam MULTI the slot needs several passes: strips exchange their edge column through `bnd` between passes.
 */
template <int FAM, int K, bool MULTI>
__device__ __forceinline__ void sweep_batch(KernelArgs const &args, SweepConstants const &sc, uint32_t const *prof,
                                            CandSlot const &slot, Job const &job, uint8_t const *c0_ptr,
                                            uint32_t batch, uint32_t n_queries, int pass, int passes, int G,
                                            uint32_t *scratch) {
    using traits = family_traits<FAM>;
    constexpr bool packed = traits::packed, affine = traits::affine, local = traits::local;
    constexpr int VECS = K / 4;
    constexpr uint32_t per_row = affine ? 2 : 1;
    int const COLS = K * G, GROUPS = 32 / G;
    int const lane = threadIdx.x & 31;
    int const gl = lane & (G - 1), gi = lane / G;
    int const open = sc.open, extend = sc.extend, bias = sc.bias, go = sc.go;
    bool const compare = FAM == FAM_LIN_COMPARE || (!packed && sc.compare);

    uint32_t const q_pos = batch * GROUPS + gi;
    bool const q_valid = q_pos < n_queries;
    uint32_t const q = q_valid ? args.q_order[job.q_begin + q_pos] : 0;
    uint32_t const q_begin_off = q_valid ? args.q_off[q] : 0;
    int const m = q_valid ? (int)(args.q_off[q + 1] - q_begin_off) : 0;
    uint8_t const *const q_ptr = args.q_cls + q_begin_off;

    // Multi-pass boundary storage: the last DP column of the previous pass, one row-slot per query row.
    // Queries of a job are sorted longest-first, so the job's first query bounds every row count.
    uint32_t *bnd = nullptr;
    if constexpr (MULTI) {
        uint32_t const first = args.q_order[job.q_begin];
        uint32_t const rows = args.q_off[first + 1] - args.q_off[first] + 1;
        bnd = scratch + (size_t)q_pos * rows * per_row;
    }

    uint32_t const pass_base = (uint32_t)pass * COLS;
    int const cbase = (int)pass_base + gl * K; // DP column left of this lane's strip
    bool const first_strip = pass == 0 && gl == 0;
    bool const from_scratch = MULTI && pass > 0 && gl == 0;
    bool const to_scratch = MULTI && pass + 1 < passes && gl == G - 1;

    // Top boundary: DP row 0 of the strip, and the seeds of the gap tracks.
    uint32_t h[K], f[affine ? K : 1];
    // Compare mode: the strip's candidate symbols — bytes (or runes when wide) in the 32-bit tier; in the packed linear
    // family both candidates' bytes, pre-shifted so that any mismatch XORs to at least 256: (c0 << 8) | (c1 << 24).
    uint32_t cb[(!packed || FAM == FAM_LIN_COMPARE) ? K : 1];
    bool const wide = !packed && args.wide_symbols != 0;
    uint32_t const *const q_wide = (uint32_t const *)args.q_cls + q_begin_off;
    uint32_t diag_in, w_out = 0, e_out = 0, best = 0;
    uint32_t left_boundary = 0, left_gap_boundary = 0; // pass 0, lane 0: DP column 0 (packed families: constant)
    if constexpr (traits::linear_shift) {
#pragma unroll
        for (int k = 0; k < K; ++k) h[k] = 0;
        diag_in = 0;
        if constexpr (FAM == FAM_LIN_COMPARE) {
            uint8_t const *const c1_ptr = slot.c1 != kNoCandidate ? args.c_cls + args.c_off[slot.c1] : c0_ptr;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                uint32_t const col = (uint32_t)cbase + k;
                // Padding halves hold 0x00FF: never equal to a shifted byte, and any XOR with one stays >= 255.
                uint32_t const low = col < slot.len0 ? (uint32_t)c0_ptr[col] << 8 : 0x00FFu;
                uint32_t const high = col < slot.len1 ? (uint32_t)c1_ptr[col] << 8 : 0x00FFu;
                cb[k] = low | (high << 16);
            }
        }
    }
    else if constexpr (FAM == FAM_AFF_SHIFT) {
        uint32_t const top = pack_bits(bias - go);
#pragma unroll
        for (int k = 0; k < K; ++k) h[k] = top, f[k] = 0;
        diag_in = first_strip ? pack_bits(bias) : top;
        left_boundary = top, left_gap_boundary = 0;
    }
    else if constexpr (FAM == FAM_SW_LIN) {
#pragma unroll
        for (int k = 0; k < K; ++k) h[k] = sc.bias_bits;
        diag_in = sc.bias_bits;
        best = sc.bias_bits;
        left_boundary = sc.bias_bits;
    }
    else if constexpr (FAM == FAM_SW_AFF) {
        // Cells are kept as H + open (what both gap tracks consume); the vertical track is kept floored at zero.
        uint32_t const zero_cell = pack_bits(bias + (SZB_SW_SHORT_CHAIN ? 0 : open));
#pragma unroll
        for (int k = 0; k < K; ++k) h[k] = zero_cell, f[k] = sc.bias_bits;
        diag_in = zero_cell;
        best = zero_cell;
        left_boundary = zero_cell, left_gap_boundary = pack_bits(bias + open + extend);
    }
    else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            int const top = (int)degenerate_score<FAM>(cbase + 1 + k, open, extend);
            h[k] = (uint32_t)top;
            if constexpr (affine) f[k] = (uint32_t)(top + open + extend);
        }
        diag_in = (uint32_t)(int)degenerate_score<FAM>(cbase, open, extend);
        if (compare) {
            uint32_t const c_begin_off = args.c_off[slot.c0];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                uint32_t const col = (uint32_t)cbase + k;
                // Padding columns hold a symbol no string contains; their cells are never read anyway.
                cb[k] = col >= slot.len0 ? 0xFFFFFFFFu
                        : wide           ? ((uint32_t const *)args.c_cls)[c_begin_off + col]
                                         : (uint32_t)c0_ptr[col];
            }
        }
    }

    // EXPERIMENT, off by default (-DSZB_SPLIT_STRIPS=1): the affine packed families split every lane's strip into two
    // sub-strips that run ONE ROW APART (sub-strip A on row r while sub-strip B finishes row r - 1), so that a lane carries
    // two independent dependency chains per step and the wavefront becomes 2 G virtual lanes wide. Measured on c3 / c3_sw /
    // c4: 6.9 / 5.2 / 5.0 TCUPS against 7.9 / 6.2 / 5.7 without it. The SASS-level profile of the single-chain kernel
    // explains why: with four warps per scheduler every dependent DPX instruction already costs a warp ~8 cycles
    // (wait + pipe throttle + not selected), i.e. the four warps together issue one DPX instruction every 2 cycles, which
    // IS the pipe's rate; the chain phase is not where the idle pipe cycles are.
    constexpr bool split = SZB_SPLIT_STRIPS != 0 && (FAM == FAM_AFF_SHIFT || FAM == FAM_SW_AFF) && K >= 16;
    constexpr int VA = split ? (VECS + 1) / 2 : VECS; // vectors of sub-strip A
    constexpr int KA = VA * 4;                        // columns of sub-strip A
    constexpr int STRIDE = split ? 2 : 1;             // virtual lanes per lane

    // Number of steps: the longest query of the warp's groups plus the pipeline fill.
    int steps = m > 0 ? m + STRIDE * G - 1 : 0;
    for (int offset = 16; offset >= G; offset >>= 1) steps = max(steps, __shfl_xor_sync(0xFFFFFFFFu, steps, offset));

    uint32_t a_next = (gl == 0 && m > 0) ? (wide ? q_wide[0] : (uint32_t)q_ptr[0]) : 0;
    bool valid_next = gl == 0 && m > 0;
    uint32_t bw_next = 0, be_next = 0;
    if constexpr (MULTI) {
        if (from_scratch && m > 0) {
            bw_next = bnd[per_row];
            if constexpr (affine) be_next = bnd[per_row + 1];
        }
    }
    // Shared-window byte address of this lane's slot in profile row 0; a row is K / 4 vectors of 32 slots x 16 bytes.
    uint32_t const strip_address = shared_window_address(prof) + (uint32_t)gl * 16u;
    constexpr uint32_t row_bytes = (uint32_t)K * 32u * 4u;

    // Sub-strip B's inputs: the row class and validity sub-strip A had one step ago, and the last cell of sub-strip A
    // after the previous step (B's left neighbour) and after the one before (B's diagonal neighbour).
    uint32_t a_b = 0, mid_w1 = split ? h[KA - 1] : 0, mid_w2 = mid_w1, mid_e1 = 0, e_b = 0;
    bool valid_b = false;

    auto add = [&](uint32_t x, uint32_t y) -> uint32_t {
        if constexpr (packed) return add_on_fma_pipe(x, y, sc.one);
        else return x + y;
    };
    // One DP cell: strip column k, diagonal sum t, left neighbour l, horizontal gap track e.
    auto cell = [&](int k, uint32_t t, uint32_t &l, uint32_t &e, uint32_t a) {
        if constexpr (traits::linear_shift) { l = h[k] = __vimax3_u16x2(t, h[k], l); }
        else if constexpr (FAM == FAM_AFF_SHIFT) {
            f[k] = __viaddmax_u16x2(h[k], sc.neg_go_bits, f[k]);
            e = __viaddmax_u16x2(l, sc.neg_go_bits, e);
            l = h[k] = __vimax3_u16x2(t, f[k], e);
        }
        else if constexpr (FAM == FAM_SW_LIN) {
            uint32_t const up_gap = add(h[k], sc.gap_arith);
            uint32_t const x = __vimax3_u16x2(t, up_gap, sc.bias_bits);
            l = h[k] = __viaddmax_u16x2(l, sc.gap_bits, x);
            if (k & 1) best = __vimax3_u16x2(best, h[k - 1], h[k]);
        }
        else if constexpr (FAM == FAM_SW_AFF) {
            // h[] holds H + open. The vertical track is floored at zero (exact: H is floored anyway and extend <= 0),
            // which leaves ONE three-way maximum for the cell: 3 DPX + 3 FMA-pipe adds (+ half a DPX for the running max).
            if constexpr (SZB_SW_SHORT_CHAIN != 0) {
                f[k] = __vimax3_u16x2(add(h[k], sc.open_arith), add(f[k], sc.ext_arith), sc.bias_bits);
                e = __viaddmax_u16x2(l, sc.gap_bits, add(e, sc.ext_arith));
                l = h[k] = __vimax3_u16x2(t, e, f[k]);
            }
            else {
                f[k] = __vimax3_u16x2(h[k], add(f[k], sc.ext_arith), sc.bias_bits);
                e = __viaddmax_u16x2(e, sc.ext_bits, l);
                l = h[k] = add(__vimax3_u16x2(t, e, f[k]), sc.open_arith);
            }
            if (k & 1) best = __vimax3_u16x2(best, h[k - 1], h[k]);
        }
        else { // signed 32-bit cells, literal recurrences
            int const ti = (int)t;
            int score;
            if constexpr (affine) {
                int const del = __viaddmax_s32((int)h[k], open, (int)f[k] + extend);
                int const ins = __viaddmax_s32((int)l, open, (int)e + extend);
                f[k] = (uint32_t)del, e = (uint32_t)ins;
                score = __vimax3_s32(ti, del, ins);
            }
            else { score = __viaddmax_s32(max((int)h[k], (int)l), open, ti); }
            if constexpr (local) {
                score = max(score, 0);
                // Gap rewards can inflate padding columns: only real cells compete.
                if (cbase + k < (int)slot.len0) best = (uint32_t)max((int)best, score);
            }
            l = h[k] = (uint32_t)score;
        }
        (void)a;
    };
    // Compare modes form the substitution term of strip column k from the strip's symbols in registers.
    auto compared = [&](int k, uint32_t a) -> uint32_t {
        if constexpr (FAM == FAM_LIN_COMPARE) return sc.cmp_match_bits - __vminu2(cb[k] ^ (a * 0x01000100u), sc.cmp_drop_bits);
        else if constexpr (!packed) return (uint32_t)(cb[k] == a ? sc.cmp_match : sc.cmp_mismatch);
        else return 0u;
    };
    // Profile rows are stored ROTATED by one column: vector v of a strip holds the terms of its columns 4v+1 .. 4v+4, the
    // last slot of the last vector the term of column 0. The diagonal sums of those columns are formed from the previous
    // row's cells the moment the vector arrives — before any of them is overwritten — so every cell of the strip is
    // updated in place and nothing has to be moved aside.
    auto first_sum = [&](int k, uint32_t diagonal, uint32_t row_address, uint32_t a) -> uint32_t {
        if (!compare) {
            // Column k's term sits in the last slot of the vector before it (column 0: of the last vector).
            uint32_t const slot_before = (uint32_t)((k == 0 ? VECS : k / 4) - 1) * 512u + 12u;
            return add(diagonal, load_shared_word(row_address + slot_before));
        }
        return add(diagonal, compared(k, a));
    };
    auto vector = [&](int v, uint32_t row_address, uint32_t a, uint32_t &t_carry, uint32_t &l, uint32_t &e) {
        uint32_t p[4];
        if (!compare) {
            load_shared_quad(row_address + (uint32_t)v * 512u, p[0], p[1], p[2], p[3]);
        }
        else {
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = compared((v * 4 + 1 + i) % K, a);
        }
        uint32_t t_next[4]; // diagonal sums of columns 4v+1 .. 4v+4 (the one past a sub-strip's end is not needed)
#pragma unroll
        for (int i = 0; i < 4; ++i) t_next[i] = add(h[v * 4 + i], p[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) cell(v * 4 + i, i == 0 ? t_carry : t_next[i - 1], l, e, a);
        t_carry = t_next[3];
    };

    for (int s = 0; s < steps; ++s) {
        // Strip's left neighbour: the last cell of the lane to the left, one step ago.
        uint32_t w_in = __shfl_up_sync(0xFFFFFFFFu, w_out, 1, G);
        uint32_t e_in = 0;
        if constexpr (affine) e_in = __shfl_up_sync(0xFFFFFFFFu, e_out, 1, G);
        if (gl == 0) {
            if (!MULTI || pass == 0) {
                if constexpr (packed) { w_in = left_boundary, e_in = left_gap_boundary; }
                else {
                    w_in = (uint32_t)(int)degenerate_score<FAM>(s + 1, open, extend); // DP row r = s + 1 on lane 0
                    if constexpr (affine) e_in = w_in + (uint32_t)(open + extend);
                }
            }
            else { w_in = bw_next, e_in = be_next; }
        }

        bool const valid = valid_next; // (sub-strip A of) this lane processes DP row s - STRIDE * gl + 1 in this step
        uint32_t const a = a_next;
        {   // Prefetch the next row's class and (multi-pass) boundary cell.
            int const next = s + 1 - STRIDE * gl;
            valid_next = (unsigned)next < (unsigned)m;
            if (valid_next) a_next = wide ? q_wide[next] : (uint32_t)q_ptr[next];
            if constexpr (MULTI) {
                if (from_scratch && valid_next) {
                    bw_next = bnd[(next + 1) * per_row];
                    if constexpr (affine) be_next = bnd[(next + 1) * per_row + 1];
                }
            }
        }

        if constexpr (!split) {
            if (valid) {
                uint32_t l = w_in, e = e_in;
                uint32_t const row_address = strip_address + a * row_bytes;
                uint32_t t_carry = first_sum(0, diag_in, row_address, a);
#pragma unroll
                for (int v = 0; v < VECS; ++v) vector(v, row_address, a, t_carry, l, e);
                diag_in = w_in;
                w_out = l, e_out = e;
                if constexpr (MULTI) {
                    if (to_scratch) {
                        int const r = s - gl + 1; // the DP row just finished
                        bnd[r * per_row] = l;
                        if constexpr (affine) bnd[r * per_row + 1] = e;
                    }
                }
            }
        }
        else {
            uint32_t const row_a = strip_address + a * row_bytes, row_b = strip_address + a_b * row_bytes;
            uint32_t l_a = w_in, e_a = e_in, l_b = mid_w1;
            e_b = mid_e1;
            if (valid && valid_b) {
                // Both sub-strips are live (all but the first and last step of a lane): one block, two chains.
                uint32_t t_a = first_sum(0, diag_in, row_a, a), t_b = first_sum(KA, mid_w2, row_b, a_b);
#pragma unroll
                for (int v = 0; v < VA; ++v) {
                    vector(v, row_a, a, t_a, l_a, e_a);
                    if (VA + v < VECS) vector(VA + v, row_b, a_b, t_b, l_b, e_b);
                }
            }
            else {
                if (valid) {
                    uint32_t t_a = first_sum(0, diag_in, row_a, a);
#pragma unroll
                    for (int v = 0; v < VA; ++v) vector(v, row_a, a, t_a, l_a, e_a);
                }
                if (valid_b) {
                    uint32_t t_b = first_sum(KA, mid_w2, row_b, a_b);
#pragma unroll
                    for (int v = VA; v < VECS; ++v) vector(v, row_b, a_b, t_b, l_b, e_b);
                }
            }
            if (valid_b) {
                w_out = l_b, e_out = e_b;
                if constexpr (MULTI) {
                    if (to_scratch) {
                        int const r = s - STRIDE * gl; // the DP row sub-strip B just finished
                        bnd[r * per_row] = l_b;
                        bnd[r * per_row + 1] = e_b;
                    }
                }
            }
            mid_w2 = mid_w1;
            if (valid) diag_in = w_in, mid_w1 = l_a, mid_e1 = e_a;
            valid_b = valid, a_b = a;
        }
    }

    // Harvest: the cell (m, n) of each half for global alignments, the running maximum for local ones.
#pragma unroll
    for (int half = 0; half < (packed ? 2 : 1); ++half) {
        uint32_t const cand = half == 0 ? slot.c0 : slot.c1;
        if (cand == kNoCandidate) continue; // uniform across the CTA
        int const n = (int)(half == 0 ? slot.len0 : slot.len1);
        int64_t *const out = args.results + (uint64_t)(q - args.row_base) * args.stride + cand;

        if constexpr (!local) {
            if (m == 0 || n == 0) {
                if (pass == 0 && gl == 0 && q_valid) {
                    int64_t const score = degenerate_score<FAM>(m == 0 ? n : m, open, extend);
                    *out = sc.negate ? -score : score;
                }
                continue;
            }
            if ((n - 1) / COLS != pass) continue; // uniform across the CTA
            int const lane_star = ((n - 1) % COLS) / K, k_star = (n - 1) % K;
            uint32_t cell = 0;
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (k == k_star) cell = h[k];
            if (gl == lane_star && q_valid) {
                int64_t score;
                if constexpr (traits::linear_shift)
                    score = (int64_t)((cell >> (16 * half)) & 0xFFFFu) + (int64_t)(m + n) * open;
                else if constexpr (FAM == FAM_AFF_SHIFT)
                    score = (int64_t)((cell >> (16 * half)) & 0xFFFFu) - bias + (int64_t)(m + n) * extend;
                else score = (int64_t)(int)cell;
                *out = sc.negate ? -score : score;
            }
        }
        else {
            // Reduce the running maximum over the group's lanes.
            uint32_t group_best = best;
            for (int offset = G / 2; offset >= 1; offset >>= 1) {
                uint32_t const other = __shfl_xor_sync(0xFFFFFFFFu, group_best, offset);
                if constexpr (packed) group_best = __vmaxu2(group_best, other);
                else group_best = (uint32_t)max((int)group_best, (int)other);
            }
            int64_t score = packed ? (int64_t)((group_best >> (16 * half)) & 0xFFFFu) - bias - ((FAM == FAM_SW_AFF && !SZB_SW_SHORT_CHAIN) ? open : 0)
                                   : (int64_t)(int)group_best;
            if (gl == 0 && q_valid) {
                // Later passes of a long candidate fold into the cell written by pass 0 (same CTA, ordered by
                // the pass barrier).
                if (m == 0 || n == 0) score = 0;
                if (pass == 0) *out = score;
                else if (score > *out) *out = score;
            }
        }
    }
}

/**
 *  @brief One persistent CTA: pulls jobs, builds profiles, sweeps queries.
 *  @tparam FAM recurrence family, @tparam K DP columns per lane (the register strip).
 */
template <int FAM, int K>
__global__ void __launch_bounds__(kThreads, 1) wavefront_kernel(KernelArgs const args) {
    using traits = family_traits<FAM>;
    constexpr bool packed = traits::packed;
    static_assert(K % 4 == 0, "strips are read as 16-byte vectors");

    SZB_DYNAMIC_SHARED(smem);
    Scoring const *const sc_global = args.scoring;
    int const n_classes = sc_global->n_classes;
    uint32_t *const ctl = smem;                // [0] job index, [1] next batch, [2..3] the mbarrier of the TMA copies
    int8_t *const sub = (int8_t *)(smem + 4);  // 1024 B
    uint8_t *const stage = (uint8_t *)(smem + 4 + 256); // 2 x kStageBytes: the slot's candidate bytes of this pass
    uint32_t *const prof = smem + 4 + 256 + 2 * kStageBytes / 4; // [n_classes][K / 4][32 slots][4]; unused in compare mode
    uint32_t const barrier = shared_window_address(ctl + 2);
    uint32_t const stage_address = shared_window_address(stage);
    uint32_t barrier_parity = 0;

    SweepConstants sc;
    sc.open = sc_global->open, sc.extend = sc_global->extend, sc.bias = sc_global->bias;
    sc.go = sc.extend - sc.open, sc.negate = sc_global->negate;
    sc.cmp_match = sc_global->match, sc.cmp_mismatch = sc_global->mismatch;
    sc.compare = FAM == FAM_LIN_COMPARE || (!packed && sc_global->compare != 0);
    {   // Byte-compare flavour of the shifted linear recurrence: e = max(S - 2g, 0) for the two possible S.
        int const e_match = max(sc.cmp_match - 2 * sc.open, 0), e_mismatch = max(sc.cmp_mismatch - 2 * sc.open, 0);
        sc.cmp_match_bits = pack_bits(e_match), sc.cmp_drop_bits = pack_bits(e_match - e_mismatch);
    }
    sc.neg_go_bits = pack_bits(-sc.go), sc.bias_bits = pack_bits(sc.bias), sc.gap_bits = pack_bits(sc.open);
    sc.gap_arith = pack_arith(sc.open, sc.open), sc.ext_arith = pack_arith(sc.extend, sc.extend);
    sc.ext_bits = pack_bits(sc.extend), sc.open_arith = pack_arith(sc.open, sc.open);
    sc.one = blockDim.x / kThreads; // 1, but not a constant ptxas can fold: the adds stay IMADs

    for (int i = threadIdx.x; i < 256; i += kThreads) ((uint32_t *)sub)[i] = ((uint32_t const *)sc_global->sub)[i];
    if (threadIdx.x == 0) {
        mbarrier_init(barrier, 1);
        mbarrier_init_fence();
    }
    int const lane = threadIdx.x & 31;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) ctl[0] = atomicAdd(args.job_counter, 1u);
        __syncthreads();
        uint32_t const job_index = ctl[0];
        if (job_index >= args.n_jobs) break;
        Job const job = args.jobs[job_index];
        CandSlot const slot = args.slots[job.slot];
        int const G = (int)job.lanes, COLS = K * G, GROUPS = 32 / G;
        uint32_t const longest = max(slot.len0, slot.len1);
        int const passes = longest == 0 ? 1 : (int)((longest + COLS - 1) / COLS);
        uint8_t const *const c0_ptr = args.c_cls + args.c_off[slot.c0];
        uint8_t const *const c1_ptr = slot.c1 != kNoCandidate ? args.c_cls + args.c_off[slot.c1] : c0_ptr;
        uint32_t *const scratch = args.scratch + (size_t)blockIdx.x * args.scratch_words_per_cta;
        uint32_t const n_queries = job.q_end - job.q_begin;
        uint32_t const n_batches = (n_queries + GROUPS - 1) / GROUPS;

        for (int pass = 0; pass < passes; ++pass) {
            uint32_t const pass_base = (uint32_t)pass * COLS;
            __syncthreads(); // every warp is done with the previous profile and batch counter
            if (threadIdx.x == 0) ctl[1] = 0;

            // 1. Build the substitution profile of this slot's columns [pass_base, pass_base + COLS). The candidate bytes
            //    of the pass arrive through the TMA engine: one thread arms the mbarrier with the byte count and issues
            //    one bulk copy per candidate (16-byte aligned windows around the columns), everybody waits on the barrier.
            if (!sc.compare) {
                uint32_t const columns0 = slot.len0 > pass_base ? min((uint32_t)COLS, slot.len0 - pass_base) : 0;
                uint32_t const columns1 = slot.len1 > pass_base ? min((uint32_t)COLS, slot.len1 - pass_base) : 0;
                uint8_t const *const begin0 = c0_ptr + pass_base, *const begin1 = c1_ptr + pass_base;
                uint32_t const skew0 = (uint32_t)((uintptr_t)begin0 & 15), skew1 = (uint32_t)((uintptr_t)begin1 & 15);
                uint32_t const bytes0 = columns0 ? (skew0 + columns0 + 15) & ~15u : 0;
                uint32_t const bytes1 = columns1 ? (skew1 + columns1 + 15) & ~15u : 0;
                if (bytes0 + bytes1) { // uniform across the CTA
                    if (threadIdx.x == 0) {
                        mbarrier_expect_bytes(barrier, bytes0 + bytes1);
                        if (bytes0) bulk_copy_to_shared(stage_address, begin0 - skew0, bytes0, barrier);
                        if (bytes1) bulk_copy_to_shared(stage_address + kStageBytes, begin1 - skew1, bytes1, barrier);
                    }
                    mbarrier_wait(barrier, barrier_parity);
                    barrier_parity ^= 1;
                }
                uint8_t const *const staged0 = stage + skew0, *const staged1 = stage + kStageBytes + skew1;
                for (int idx = threadIdx.x; idx < n_classes * COLS; idx += kThreads) {
                    int const a = idx / COLS, j = idx - a * COLS;
                    bool const v0 = (uint32_t)j < columns0, v1 = (uint32_t)j < columns1;
                    int const b0 = v0 ? staged0[j] : 0, b1 = v1 ? staged1[j] : 0;
                    int const e0 = profile_term<FAM>(sub, a, b0, v0, sc.open, sc.extend, sc.bias);
                    int const e1 = profile_term<FAM>(sub, a, b1, v1, sc.open, sc.extend, sc.bias);
                    // Row layout: [vector][strip slot, 32 per vector][4 words]; a strip is one lane's K columns, so lane
                    // `gl` reads its v-th vector at the compile-time offset v * 512 bytes.
                    int const strip = j / K, k = j - strip * K;
                    int const rotated = k == 0 ? K - 1 : k - 1; // column k's term sits one slot early (see sweep_batch)
                    int const word = ((rotated / 4) * 32 + strip) * 4 + (rotated % 4);
                    prof[a * (K * 32) + word] = packed ? pack_arith(e0, e1) : (uint32_t)e0;
                }
            }
            __syncthreads();

            // 2. Warps pull batches of GROUPS queries and sweep them over the profile.
            for (;;) {
                uint32_t batch = 0;
                if (lane == 0) batch = atomicAdd(&ctl[1], 1u);
                batch = __shfl_sync(0xFFFFFFFFu, batch, 0);
                if (batch >= n_batches) break;
                if (passes > 1)
                    sweep_batch<FAM, K, true>(args, sc, prof, slot, job, c0_ptr, batch, n_queries, pass, passes, G,
                                              scratch);
                else
                    sweep_batch<FAM, K, false>(args, sc, prof, slot, job, c0_ptr, batch, n_queries, 0, 1, G, scratch);
            }
        }
    }
}

} // namespace szb
