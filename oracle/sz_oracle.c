/*
 * oracle/sz_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into, imported by, or called from the product path).
 *
 * A plain-C, CPU restatement of the reference's batched string-similarity semantics, used as the parity checker for
 * the CUDA path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may load this.
 *
 * Every function restates the *semantics* pinned by the reference's own "second opinion" dual-row baselines and its
 * serial engine, and is pinned (tests/test_oracle.py) against:
 *   - the 15 known-answer Levenshtein vectors of /root/reference/test/similarities.cuh:609-625 (tests/golden/),
 *   - golden matrices generated from the UNMODIFIED reference CPU library (oracle/_ref, built by oracle/build_ref.sh)
 *     with serial and Ice Lake capability masks (tests/golden/make_golden.py -> tests/golden/*.json),
 *   - closed forms of test/similarities.cuh:1042-1058 (d(x,"") = |x|*gap, d(x,x) = 0).
 * Parity is therefore PINNED.
 *
 * Reference citations (relative to /root/reference):
 *   linear Levenshtein .......... test/similarities.cuh:47-70, include/stringzillas/similarities/serial.hpp:846-849
 *   linear Needleman-Wunsch ..... test/similarities.cuh:73-100
 *   linear Smith-Waterman ....... test/similarities.cuh:103-133, serial.hpp:957-965
 *   affine (Gotoh) Levenshtein .. test/similarities.cuh:136-180, serial.hpp:1049-1056,1091-1102
 *   affine Needleman-Wunsch ..... test/similarities.cuh:186-228
 *   affine Smith-Waterman ....... test/similarities.cuh:234-280, serial.hpp:1195-1201,1231-1247
 *   open == extend -> linear .... c/stringzillas/levenshtein.cuh:117, c/stringzillas/needleman_wunsch.cuh:107
 *   cross product / symmetric ... serial.hpp:3140-3187 (lower triangle incl. diagonal, scored as (query=i, cand=j<=i)
 *                                  and mirrored), include/stringzillas/stringzillas.h:173-177
 *   substitution lookup ......... serial.hpp:193-200: cost = table[byte_to_class[q]][byte_to_class[c]] (query = row)
 *   UTF-8 Levenshtein ........... serial.hpp:2796-2898: both strings are transcoded to UTF-32 with the UNCHECKED decoder
 *                                  (include/stringzilla/utf8_runes/serial.h:111-124) and the same recurrences run over runes
 *
 * Arithmetic: Levenshtein accumulates in unsigned 64-bit exactly like the baselines' std::size_t (the int8 costs are
 * converted as the C++ baselines convert them); NW/SW accumulate in signed 64-bit (std::ptrdiff_t).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

static inline uint64_t min_u64(uint64_t a, uint64_t b) { return a < b ? a : b; }
static inline int64_t max_i64(int64_t a, int64_t b) { return a > b ? a : b; }

/* test/similarities.cuh:47-70 */
static uint64_t levenshtein_linear(uint8_t const *s1, size_t len1, uint8_t const *s2, size_t len2, int8_t match,
                                   int8_t mismatch, int8_t gap, uint64_t *scratch) {
    size_t const cols = len2 + 1;
    uint64_t *previous = scratch, *current = scratch + cols;
    for (size_t j = 0; j < cols; ++j) previous[j] = (uint64_t)j * (uint64_t)(int64_t)gap;
    for (size_t i = 1; i <= len1; ++i) {
        current[0] = (uint64_t)i * (uint64_t)(int64_t)gap;
        uint8_t const a = s1[i - 1];
        for (size_t j = 1; j < cols; ++j) {
            uint64_t substitution = (uint64_t)(int64_t)(a == s2[j - 1] ? match : mismatch);
            uint64_t if_gap = min_u64(previous[j], current[j - 1]) + (uint64_t)(int64_t)gap;
            current[j] = min_u64(if_gap, previous[j - 1] + substitution);
        }
        uint64_t *t = previous;
        previous = current, current = t;
    }
    return previous[cols - 1];
}

/* test/similarities.cuh:136-180 */
static uint64_t levenshtein_affine(uint8_t const *s1, size_t len1, uint8_t const *s2, size_t len2, int8_t match,
                                   int8_t mismatch, int8_t open, int8_t extend, uint64_t *scratch) {
    size_t const cols = len2 + 1;
    uint64_t *ps = scratch, *cs = scratch + cols, *pi = scratch + 2 * cols, *ci = scratch + 3 * cols,
             *pd = scratch + 4 * cols, *cd = scratch + 5 * cols;
    uint64_t const o = (uint64_t)(int64_t)open, e = (uint64_t)(int64_t)extend;
    memset(scratch, 0, sizeof(uint64_t) * 6 * cols);
    ps[0] = 0;
    for (size_t j = 1; j < cols; ++j) ps[j] = o + (uint64_t)(j - 1) * e, pd[j] = ps[j] + o + e;
    for (size_t i = 1; i <= len1; ++i) {
        cs[0] = o + (uint64_t)(i - 1) * e;
        ci[0] = cs[0] + o + e;
        uint8_t const a = s1[i - 1];
        for (size_t j = 1; j < cols; ++j) {
            uint64_t substitution = (uint64_t)(int64_t)(a == s2[j - 1] ? match : mismatch);
            uint64_t if_substitution = ps[j - 1] + substitution;
            uint64_t if_insertion = min_u64(cs[j - 1] + o, ci[j - 1] + e);
            uint64_t if_deletion = min_u64(ps[j] + o, pd[j] + e);
            cs[j] = min_u64(min_u64(if_deletion, if_insertion), if_substitution);
            ci[j] = if_insertion, cd[j] = if_deletion;
        }
        uint64_t *t;
        t = ps, ps = cs, cs = t;
        t = pi, pi = ci, ci = t;
        t = pd, pd = cd, cd = t;
    }
    return ps[cols - 1];
}

/* test/similarities.cuh:73-133 — `local` switches Needleman-Wunsch (global) to Smith-Waterman. */
static int64_t weighted_linear(uint8_t const *s1, size_t len1, uint8_t const *s2, size_t len2,
                               uint8_t const *byte_to_class, int8_t const *costs, int8_t gap, int local,
                               int64_t *scratch) {
    size_t const cols = len2 + 1;
    int64_t *previous = scratch, *current = scratch + cols;
    int64_t best = 0;
    for (size_t j = 0; j < cols; ++j) previous[j] = local ? 0 : (int64_t)j * gap;
    for (size_t i = 1; i <= len1; ++i) {
        current[0] = local ? 0 : (int64_t)i * gap;
        int8_t const *row = costs + 32 * (size_t)byte_to_class[s1[i - 1]];
        for (size_t j = 1; j < cols; ++j) {
            int64_t if_substitution = previous[j - 1] + row[byte_to_class[s2[j - 1]]];
            int64_t if_gap = max_i64(previous[j], current[j - 1]) + gap;
            if (local) if_substitution = max_i64(if_substitution, 0);
            int64_t score = max_i64(if_gap, if_substitution);
            current[j] = score;
            best = max_i64(best, score);
        }
        int64_t *t = previous;
        previous = current, current = t;
    }
    return local ? best : previous[cols - 1];
}

/* test/similarities.cuh:186-280 */
static int64_t weighted_affine(uint8_t const *s1, size_t len1, uint8_t const *s2, size_t len2,
                               uint8_t const *byte_to_class, int8_t const *costs, int8_t open, int8_t extend, int local,
                               int64_t *scratch) {
    size_t const cols = len2 + 1;
    int64_t *ps = scratch, *cs = scratch + cols, *pi = scratch + 2 * cols, *ci = scratch + 3 * cols,
            *pd = scratch + 4 * cols, *cd = scratch + 5 * cols;
    int64_t best = 0;
    memset(scratch, 0, sizeof(int64_t) * 6 * cols);
    ps[0] = 0;
    for (size_t j = 1; j < cols; ++j) {
        ps[j] = local ? 0 : (int64_t)open + (int64_t)(j - 1) * extend;
        pd[j] = ps[j] + open + extend;
    }
    for (size_t i = 1; i <= len1; ++i) {
        cs[0] = local ? 0 : (int64_t)open + (int64_t)(i - 1) * extend;
        ci[0] = cs[0] + open + extend;
        int8_t const *row = costs + 32 * (size_t)byte_to_class[s1[i - 1]];
        for (size_t j = 1; j < cols; ++j) {
            int64_t if_substitution = ps[j - 1] + row[byte_to_class[s2[j - 1]]];
            int64_t if_insertion = max_i64(cs[j - 1] + open, ci[j - 1] + extend);
            int64_t if_deletion = max_i64(ps[j] + open, pd[j] + extend);
            if (local) if_substitution = max_i64(if_substitution, 0);
            int64_t score = max_i64(max_i64(if_deletion, if_insertion), if_substitution);
            cs[j] = score, ci[j] = if_insertion, cd[j] = if_deletion;
            best = max_i64(best, score);
        }
        int64_t *t;
        t = ps, ps = cs, cs = t;
        t = pi, pi = ci, ci = t;
        t = pd, pd = cd, cd = t;
    }
    return local ? best : ps[cols - 1];
}

/* Single-pair entry points (open == extend selects the linear recurrence, levenshtein.cuh:117). */
ORACLE_API uint64_t oracle_levenshtein(uint8_t const *a, size_t la, uint8_t const *b, size_t lb, int8_t match,
                                       int8_t mismatch, int8_t open, int8_t extend) {
    uint64_t *scratch = (uint64_t *)malloc(sizeof(uint64_t) * 6 * (lb + 1));
    uint64_t r = open == extend ? levenshtein_linear(a, la, b, lb, match, mismatch, open, scratch)
                                : levenshtein_affine(a, la, b, lb, match, mismatch, open, extend, scratch);
    free(scratch);
    return r;
}

ORACLE_API int64_t oracle_weighted(uint8_t const *a, size_t la, uint8_t const *b, size_t lb,
                                   uint8_t const *byte_to_class, int8_t const *costs, int8_t open, int8_t extend,
                                   int local) {
    int64_t *scratch = (int64_t *)malloc(sizeof(int64_t) * 6 * (lb + 1));
    int64_t r = open == extend ? weighted_linear(a, la, b, lb, byte_to_class, costs, open, local, scratch)
                               : weighted_affine(a, la, b, lb, byte_to_class, costs, open, extend, local, scratch);
    free(scratch);
    return r;
}

/*
 * Cross-product drivers over Arrow-style tapes with 64-bit offsets (stringzillas.h:77-92). `c_data == NULL` requests
 * the symmetric self-similarity of the queries: lower triangle including the diagonal, mirrored (serial.hpp:3171-3185).
 * OpenMP over rows is only a convenience to keep the CPU test-suite within minutes; the arithmetic is per cell.
 */
ORACLE_API int oracle_levenshtein_cross(uint8_t const *q_data, uint64_t const *q_off, size_t q_count,
                                        uint8_t const *c_data, uint64_t const *c_off, size_t c_count, int8_t match,
                                        int8_t mismatch, int8_t open, int8_t extend, uint64_t *results,
                                        size_t row_stride) {
    int const symmetric = c_data == NULL;
    if (symmetric) c_data = q_data, c_off = q_off, c_count = q_count;
    size_t longest = 0;
    for (size_t j = 0; j < c_count; ++j)
        if (c_off[j + 1] - c_off[j] > longest) longest = c_off[j + 1] - c_off[j];
#pragma omp parallel
    {
        uint64_t *scratch = (uint64_t *)malloc(sizeof(uint64_t) * 6 * (longest + 1));
#pragma omp for collapse(2) schedule(dynamic, 1)
        for (long long i = 0; i < (long long)q_count; ++i) {
            for (long long j = 0; j < (long long)c_count; ++j) {
                if (symmetric && j > i) continue; /* lower triangle incl. the diagonal, mirrored */
                uint8_t const *a = q_data + q_off[i], *b = c_data + c_off[j];
                size_t la = q_off[i + 1] - q_off[i], lb = c_off[j + 1] - c_off[j];
                uint64_t r = open == extend ? levenshtein_linear(a, la, b, lb, match, mismatch, open, scratch)
                                            : levenshtein_affine(a, la, b, lb, match, mismatch, open, extend, scratch);
                results[(size_t)i * row_stride + (size_t)j] = r;
                if (symmetric && j != i) results[(size_t)j * row_stride + (size_t)i] = r;
            }
        }
        free(scratch);
    }
    return 0;
}

ORACLE_API int oracle_weighted_cross(uint8_t const *q_data, uint64_t const *q_off, size_t q_count,
                                     uint8_t const *c_data, uint64_t const *c_off, size_t c_count,
                                     uint8_t const *byte_to_class, int8_t const *costs, int8_t open, int8_t extend,
                                     int local, int64_t *results, size_t row_stride) {
    int const symmetric = c_data == NULL;
    if (symmetric) c_data = q_data, c_off = q_off, c_count = q_count;
    size_t longest = 0;
    for (size_t j = 0; j < c_count; ++j)
        if (c_off[j + 1] - c_off[j] > longest) longest = c_off[j + 1] - c_off[j];
#pragma omp parallel
    {
        int64_t *scratch = (int64_t *)malloc(sizeof(int64_t) * 6 * (longest + 1));
#pragma omp for collapse(2) schedule(dynamic, 1)
        for (long long i = 0; i < (long long)q_count; ++i) {
            for (long long j = 0; j < (long long)c_count; ++j) {
                if (symmetric && j > i) continue;
                uint8_t const *a = q_data + q_off[i], *b = c_data + c_off[j];
                size_t la = q_off[i + 1] - q_off[i], lb = c_off[j + 1] - c_off[j];
                int64_t r = open == extend
                                ? weighted_linear(a, la, b, lb, byte_to_class, costs, open, local, scratch)
                                : weighted_affine(a, la, b, lb, byte_to_class, costs, open, extend, local, scratch);
                results[(size_t)i * row_stride + (size_t)j] = r;
                if (symmetric && j != i) results[(size_t)j * row_stride + (size_t)i] = r;
            }
        }
        free(scratch);
    }
    return 0;
}

/* ---- UTF-8 Levenshtein: the same recurrences over runes -------------------------------------------------------------- */

/* include/stringzilla/utf8_runes/serial.h:111-124 — the lead byte alone decides the length; no validation. Bytes past
 * the end read as zero here (the reference over-reads; well-formed input never gets there). */
static size_t decode_utf8(uint8_t const *text, size_t length, uint32_t *runes) {
    size_t produced = 0;
    for (size_t i = 0; i < length;) {
        uint32_t lead = text[i];
        size_t n = 1 + (lead >= 0xC0U) + (lead >= 0xE0U) + (lead >= 0xF0U);
        uint32_t b1 = i + 1 < length ? text[i + 1] : 0, b2 = i + 2 < length ? text[i + 2] : 0,
                 b3 = i + 3 < length ? text[i + 3] : 0;
        uint32_t rune;
        if (n == 1) rune = lead;
        else if (n == 2) rune = (lead & 0x1FU) << 6 | (b1 & 0x3FU);
        else if (n == 3) rune = (lead & 0x0FU) << 12 | (b1 & 0x3FU) << 6 | (b2 & 0x3FU);
        else rune = (lead & 0x07U) << 18 | (b1 & 0x3FU) << 12 | (b2 & 0x3FU) << 6 | (b3 & 0x3FU);
        runes[produced++] = rune;
        i += n;
    }
    return produced;
}

static uint64_t levenshtein_runes(uint32_t const *s1, size_t len1, uint32_t const *s2, size_t len2, int8_t match,
                                  int8_t mismatch, int8_t open, int8_t extend, uint64_t *scratch) {
    size_t const cols = len2 + 1;
    uint64_t const o = (uint64_t)(int64_t)open, e = (uint64_t)(int64_t)extend;
    if (open == extend) { /* test/similarities.cuh:47-70 over runes */
        uint64_t *previous = scratch, *current = scratch + cols;
        for (size_t j = 0; j < cols; ++j) previous[j] = (uint64_t)j * o;
        for (size_t i = 1; i <= len1; ++i) {
            current[0] = (uint64_t)i * o;
            for (size_t j = 1; j < cols; ++j) {
                uint64_t substitution = (uint64_t)(int64_t)(s1[i - 1] == s2[j - 1] ? match : mismatch);
                current[j] = min_u64(min_u64(previous[j], current[j - 1]) + o, previous[j - 1] + substitution);
            }
            uint64_t *t = previous;
            previous = current, current = t;
        }
        return previous[cols - 1];
    }
    /* test/similarities.cuh:136-180 over runes */
    uint64_t *ps = scratch, *cs = scratch + cols, *pi = scratch + 2 * cols, *ci = scratch + 3 * cols,
             *pd = scratch + 4 * cols, *cd = scratch + 5 * cols;
    memset(scratch, 0, sizeof(uint64_t) * 6 * cols);
    for (size_t j = 1; j < cols; ++j) ps[j] = o + (uint64_t)(j - 1) * e, pd[j] = ps[j] + o + e;
    for (size_t i = 1; i <= len1; ++i) {
        cs[0] = o + (uint64_t)(i - 1) * e;
        ci[0] = cs[0] + o + e;
        for (size_t j = 1; j < cols; ++j) {
            uint64_t substitution = (uint64_t)(int64_t)(s1[i - 1] == s2[j - 1] ? match : mismatch);
            uint64_t if_insertion = min_u64(cs[j - 1] + o, ci[j - 1] + e);
            uint64_t if_deletion = min_u64(ps[j] + o, pd[j] + e);
            cs[j] = min_u64(min_u64(if_deletion, if_insertion), ps[j - 1] + substitution);
            ci[j] = if_insertion, cd[j] = if_deletion;
        }
        uint64_t *t;
        t = ps, ps = cs, cs = t;
        t = pi, pi = ci, ci = t;
        t = pd, pd = cd, cd = t;
    }
    return ps[cols - 1];
}

ORACLE_API int oracle_levenshtein_utf8_cross(uint8_t const *q_data, uint64_t const *q_off, size_t q_count,
                                             uint8_t const *c_data, uint64_t const *c_off, size_t c_count,
                                             int8_t match, int8_t mismatch, int8_t open, int8_t extend,
                                             uint64_t *results, size_t row_stride) {
    int const symmetric = c_data == NULL;
    if (symmetric) c_data = q_data, c_off = q_off, c_count = q_count;
    size_t longest_q = 0, longest_c = 0;
    for (size_t i = 0; i < q_count; ++i)
        if (q_off[i + 1] - q_off[i] > longest_q) longest_q = q_off[i + 1] - q_off[i];
    for (size_t j = 0; j < c_count; ++j)
        if (c_off[j + 1] - c_off[j] > longest_c) longest_c = c_off[j + 1] - c_off[j];
#pragma omp parallel
    {
        uint64_t *scratch = (uint64_t *)malloc(sizeof(uint64_t) * 6 * (longest_c + 1));
        uint32_t *qr = (uint32_t *)malloc(sizeof(uint32_t) * (longest_q + 1));
        uint32_t *cr = (uint32_t *)malloc(sizeof(uint32_t) * (longest_c + 1));
#pragma omp for schedule(dynamic, 1)
        for (long long i = 0; i < (long long)q_count; ++i) {
            size_t const nq = decode_utf8(q_data + q_off[i], q_off[i + 1] - q_off[i], qr);
            size_t const end = symmetric ? (size_t)i + 1 : c_count;
            for (size_t j = 0; j < end; ++j) {
                size_t const nc = decode_utf8(c_data + c_off[j], c_off[j + 1] - c_off[j], cr);
                uint64_t r = levenshtein_runes(qr, nq, cr, nc, match, mismatch, open, extend, scratch);
                results[(size_t)i * row_stride + j] = r;
                if (symmetric && j != (size_t)i) results[j * row_stride + (size_t)i] = r;
            }
        }
        free(scratch), free(qr), free(cr);
    }
    return 0;
}
