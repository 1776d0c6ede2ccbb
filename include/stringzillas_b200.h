/**
 *  @file include/stringzillas_b200.h
 *  @brief C ABI of `libstringzillas_b200.so` — the B200 (sm_100a) drop-in for the batched string-similarity hot path of
 *         StringZilla's `stringzillas` layer.
 *
 *  The library exports the SAME symbol names, argument order and status codes as the reference's
 *  `include/stringzillas/stringzillas.h` (StringZilla 5.1.2), so the reference's existing C / C++ / Python / Rust
 *  bindings link against it unchanged. Each declaration below cites the reference interface it replaces
 *  (paths relative to the reference tree). The ABI carries only plain pointers, sizes and POD structs.
 *
 *  Semantics shared by every `szs_*_distances*` / `szs_*_scores*` call (reference stringzillas.h:173-177):
 *   - every query is scored against every candidate, `results[q * results_row_stride + c]`;
 *   - `candidates == NULL` requests the symmetric self-similarity of `queries` (lower triangle incl. the diagonal is
 *     scored as (query = i, candidate = j <= i) and mirrored);
 *   - `results_row_stride` is in ELEMENTS (reference c/stringzillas/levenshtein.cuh:44-45);
 *   - calls are synchronous: results are host-readable on return;
 *   - the engine handle must be NULL on entry to `*_init` and is released by `*_free`; engines own grow-only device
 *     scratch and are not re-entrant (one call at a time per engine).
 *
 *  Differences from the reference CUDA backend (all relaxations, none observable through the reference's tests):
 *   - inputs may live in pageable/pinned HOST memory, managed memory or device memory; the library stages them itself
 *     (the reference returns `sz_device_memory_mismatch_k` for plain host strings, cuda.cuh:4268-4272);
 *   - there is NO CPU fallback: with a CPU-only scope, or when no sm_100-class GPU is present, calls fail with
 *     `sz_device_code_mismatch_k` / `sz_missing_gpu_k`;
 *   - one extra, optional entry point drives the multi-GPU partition (`szs_b200_*`, bottom of this file).
 */
#ifndef STRINGZILLAS_B200_H_
#define STRINGZILLAS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SZS_B200_API __attribute__((visibility("default")))
#else
#define SZS_B200_API
#endif

/* ---- Plain types (reference include/stringzilla/types.h) ------------------------------------------------------ */

typedef uint64_t sz_size_t;      /* types.h: pointer-sized unsigned; this ABI is 64-bit only */
typedef int64_t sz_ssize_t;      /* types.h: pointer-sized signed */
typedef char const *sz_cptr_t;   /* types.h:736 */
typedef int8_t sz_error_cost_t;  /* types.h:737 */
typedef uint8_t sz_u8_t;
typedef uint32_t sz_u32_t;
typedef uint64_t sz_u64_t;
typedef sz_size_t sz_sorted_idx_t;

/** Status codes, numerically identical to reference types.h:810-833. */
typedef enum sz_status_t {
    sz_success_k = 0,
    sz_bad_alloc_k = -10,
    sz_invalid_utf8_k = -12,
    sz_contains_duplicates_k = -13,
    sz_overflow_risk_k = -14,
    sz_unexpected_dimensions_k = -15,
    sz_missing_gpu_k = -16,
    sz_device_code_mismatch_k = -17,
    sz_device_memory_mismatch_k = -18,
    sz_authentication_failed_k = -19,
    sz_status_unknown_k = -1
} sz_status_t;

/** Capability bits, numerically identical to reference types.h:839-888 (only the ones this library reads or reports). */
typedef enum sz_capability_t {
    sz_cap_serial_k = 1,
    sz_cap_parallel_k = 1 << 2,
    sz_cap_any_k = 0x7FFFFFFF,
    sz_cap_cuda_k = 1 << 21,
    sz_cap_kepler_k = 1 << 22,
    sz_cap_hopper_k = 1 << 23,
    sz_caps_none_k = 0,
    sz_caps_ck_k = (1 << 21) | (1 << 22),
    sz_caps_ckh_k = (1 << 21) | (1 << 22) | (1 << 23),
    sz_caps_cuda_k = (1 << 21) | (1 << 22) | (1 << 23)
} sz_capability_t;

/** Allocator triple, reference types.h:1007-1019. Accepted and ignored, as in the reference (levenshtein.cuh:112). */
typedef void *(*sz_memory_allocate_t)(sz_size_t, void *);
typedef void (*sz_memory_free_t)(void *, sz_size_t, void *);
typedef struct sz_memory_allocator_t {
    sz_memory_allocate_t allocate;
    sz_memory_free_t free;
    void *handle;
} sz_memory_allocator_t;

/** Callback-based string collection, reference types.h:1381-1398. */
typedef sz_cptr_t (*sz_sequence_member_start_t)(void const *, sz_sorted_idx_t);
typedef sz_size_t (*sz_sequence_member_length_t)(void const *, sz_sorted_idx_t);
typedef struct sz_sequence_t {
    void const *handle;
    sz_size_t count;
    sz_sequence_member_start_t get_start;
    sz_sequence_member_length_t get_length;
} sz_sequence_t;

/** Arrow-style tapes with N+1 offsets for N strings, reference stringzillas.h:77-92. */
typedef struct sz_sequence_u32tape_t {
    sz_cptr_t data;
    sz_u32_t const *offsets;
    sz_size_t count;
} sz_sequence_u32tape_t;
typedef struct sz_sequence_u64tape_t {
    sz_cptr_t data;
    sz_u64_t const *offsets;
    sz_size_t count;
} sz_sequence_u64tape_t;

/* ---- Library metadata (reference stringzillas.h:36-71, c/stringzillas/runtime.cuh:19-56) ----------------------- */

SZS_B200_API int szs_version_major(void);
SZS_B200_API int szs_version_minor(void);
SZS_B200_API int szs_version_patch(void);
/** What this binary ships: serial host glue + the CUDA tiers (`sz_caps_cuda_k`). */
SZS_B200_API sz_capability_t szs_capabilities_comptime(void);
/** What this machine offers: GPU bits only when device 0 is sm_100-class or newer. A missing driver yields 0 GPU bits. */
SZS_B200_API sz_capability_t szs_capabilities_runtime(void);
SZS_B200_API sz_capability_t szs_capabilities(void);

/* ---- Unified memory helpers (reference stringzillas.h:94-100,600-613, runtime.cuh:205-221) -------------------- */

SZS_B200_API sz_status_t sz_memory_allocator_init_unified(sz_memory_allocator_t *alloc, char const **error_message);
SZS_B200_API void *szs_unified_alloc(sz_size_t size_bytes);
SZS_B200_API void szs_unified_free(void *ptr, sz_size_t size_bytes);

/* ---- Device scopes (reference stringzillas.h:102-171, runtime.cuh:74-199) ------------------------------------- */

typedef void *szs_device_scope_t;
SZS_B200_API sz_status_t szs_device_scope_init_default(szs_device_scope_t *scope, char const **error_message);
SZS_B200_API sz_status_t szs_device_scope_init_cpu_cores(sz_size_t cpu_cores, szs_device_scope_t *scope,
                                                         char const **error_message);
SZS_B200_API sz_status_t szs_device_scope_init_gpu_device(sz_size_t gpu_device, szs_device_scope_t *scope,
                                                          char const **error_message);
SZS_B200_API sz_status_t szs_device_scope_get_cpu_cores(szs_device_scope_t scope, sz_size_t *cpu_cores,
                                                        char const **error_message);
SZS_B200_API sz_status_t szs_device_scope_get_gpu_device(szs_device_scope_t scope, sz_size_t *gpu_device,
                                                         char const **error_message);
SZS_B200_API sz_status_t szs_device_scope_get_capabilities(szs_device_scope_t scope, sz_capability_t *capabilities,
                                                           char const **error_message);
SZS_B200_API void szs_device_scope_free(szs_device_scope_t scope);

/* ---- Levenshtein distances (reference stringzillas.h:178-254, c/stringzillas/levenshtein.cuh:107-245) --------- */

typedef void *szs_levenshtein_distances_t;
/** `open == extend` selects linear gaps (levenshtein.cuh:117); a gap of length L costs `open + (L-1) * extend`. */
SZS_B200_API sz_status_t szs_levenshtein_distances_init(sz_error_cost_t match, sz_error_cost_t mismatch,
                                                        sz_error_cost_t open, sz_error_cost_t extend,
                                                        sz_memory_allocator_t const *alloc,
                                                        sz_capability_t capabilities,
                                                        szs_levenshtein_distances_t *engine,
                                                        char const **error_message);
SZS_B200_API sz_status_t szs_levenshtein_distances(szs_levenshtein_distances_t engine, szs_device_scope_t device,
                                                   sz_sequence_t const *queries, sz_sequence_t const *candidates,
                                                   sz_size_t *results, sz_size_t results_row_stride,
                                                   char const **error_message);
SZS_B200_API sz_status_t szs_levenshtein_distances_u32tape(szs_levenshtein_distances_t engine,
                                                           szs_device_scope_t device,
                                                           sz_sequence_u32tape_t const *queries,
                                                           sz_sequence_u32tape_t const *candidates,
                                                           sz_size_t *results, sz_size_t results_row_stride,
                                                           char const **error_message);
SZS_B200_API sz_status_t szs_levenshtein_distances_u64tape(szs_levenshtein_distances_t engine,
                                                           szs_device_scope_t device,
                                                           sz_sequence_u64tape_t const *queries,
                                                           sz_sequence_u64tape_t const *candidates,
                                                           sz_size_t *results, sz_size_t results_row_stride,
                                                           char const **error_message);
SZS_B200_API void szs_levenshtein_distances_free(szs_levenshtein_distances_t engine);

/* ---- UTF-8 Levenshtein (reference stringzillas.h:256-328, c/stringzillas/levenshtein.cuh:247-371): edit distance over
 *      UTF-8 runes. Strings are decoded with the reference's unchecked rule (utf8_runes/serial.h:111-124); pure-ASCII
 *      input takes the byte-level kernels, anything else the 32-bit-symbol compare recurrences. Unlike the reference's
 *      CUDA backend, custom and affine gap costs are supported. --------------------------------------------------- */

typedef void *szs_levenshtein_distances_utf8_t;
SZS_B200_API sz_status_t szs_levenshtein_distances_utf8_init(sz_error_cost_t match, sz_error_cost_t mismatch,
                                                             sz_error_cost_t open, sz_error_cost_t extend,
                                                             sz_memory_allocator_t const *alloc,
                                                             sz_capability_t capabilities,
                                                             szs_levenshtein_distances_utf8_t *engine,
                                                             char const **error_message);
SZS_B200_API sz_status_t szs_levenshtein_distances_utf8(szs_levenshtein_distances_utf8_t engine,
                                                        szs_device_scope_t device, sz_sequence_t const *queries,
                                                        sz_sequence_t const *candidates, sz_size_t *results,
                                                        sz_size_t results_row_stride, char const **error_message);
SZS_B200_API sz_status_t szs_levenshtein_distances_utf8_u32tape(szs_levenshtein_distances_utf8_t engine,
                                                                szs_device_scope_t device,
                                                                sz_sequence_u32tape_t const *queries,
                                                                sz_sequence_u32tape_t const *candidates,
                                                                sz_size_t *results, sz_size_t results_row_stride,
                                                                char const **error_message);
SZS_B200_API sz_status_t szs_levenshtein_distances_utf8_u64tape(szs_levenshtein_distances_utf8_t engine,
                                                                szs_device_scope_t device,
                                                                sz_sequence_u64tape_t const *queries,
                                                                sz_sequence_u64tape_t const *candidates,
                                                                sz_size_t *results, sz_size_t results_row_stride,
                                                                char const **error_message);
SZS_B200_API void szs_levenshtein_distances_utf8_free(szs_levenshtein_distances_utf8_t engine);

/* ---- Needleman-Wunsch global scores (reference stringzillas.h:337-413, c/stringzillas/needleman_wunsch.cuh) ---- */

typedef void *szs_needleman_wunsch_scores_t;
/** `byte_to_class`: 256 bytes -> one of 32 classes; `class_substitution_costs`: row-major 32x32 int8, indexed
 *  [class(query byte)][class(candidate byte)]; gaps are ADDED to a MAXIMISED score, so pass negatives. */
SZS_B200_API sz_status_t szs_needleman_wunsch_scores_init(sz_u8_t const *byte_to_class,
                                                          sz_error_cost_t const *class_substitution_costs,
                                                          sz_error_cost_t open, sz_error_cost_t extend,
                                                          sz_memory_allocator_t const *alloc,
                                                          sz_capability_t capabilities,
                                                          szs_needleman_wunsch_scores_t *engine,
                                                          char const **error_message);
SZS_B200_API sz_status_t szs_needleman_wunsch_scores(szs_needleman_wunsch_scores_t engine, szs_device_scope_t device,
                                                     sz_sequence_t const *queries, sz_sequence_t const *candidates,
                                                     sz_ssize_t *results, sz_size_t results_row_stride,
                                                     char const **error_message);
SZS_B200_API sz_status_t szs_needleman_wunsch_scores_u32tape(szs_needleman_wunsch_scores_t engine,
                                                             szs_device_scope_t device,
                                                             sz_sequence_u32tape_t const *queries,
                                                             sz_sequence_u32tape_t const *candidates,
                                                             sz_ssize_t *results, sz_size_t results_row_stride,
                                                             char const **error_message);
SZS_B200_API sz_status_t szs_needleman_wunsch_scores_u64tape(szs_needleman_wunsch_scores_t engine,
                                                             szs_device_scope_t device,
                                                             sz_sequence_u64tape_t const *queries,
                                                             sz_sequence_u64tape_t const *candidates,
                                                             sz_ssize_t *results, sz_size_t results_row_stride,
                                                             char const **error_message);
SZS_B200_API void szs_needleman_wunsch_scores_free(szs_needleman_wunsch_scores_t engine);

/* ---- Smith-Waterman local scores (reference stringzillas.h:415-488, c/stringzillas/smith_waterman.cuh) --------- */

typedef void *szs_smith_waterman_scores_t;
SZS_B200_API sz_status_t szs_smith_waterman_scores_init(sz_u8_t const *byte_to_class,
                                                        sz_error_cost_t const *class_substitution_costs,
                                                        sz_error_cost_t open, sz_error_cost_t extend,
                                                        sz_memory_allocator_t const *alloc,
                                                        sz_capability_t capabilities,
                                                        szs_smith_waterman_scores_t *engine,
                                                        char const **error_message);
SZS_B200_API sz_status_t szs_smith_waterman_scores(szs_smith_waterman_scores_t engine, szs_device_scope_t device,
                                                   sz_sequence_t const *queries, sz_sequence_t const *candidates,
                                                   sz_ssize_t *results, sz_size_t results_row_stride,
                                                   char const **error_message);
SZS_B200_API sz_status_t szs_smith_waterman_scores_u32tape(szs_smith_waterman_scores_t engine,
                                                           szs_device_scope_t device,
                                                           sz_sequence_u32tape_t const *queries,
                                                           sz_sequence_u32tape_t const *candidates,
                                                           sz_ssize_t *results, sz_size_t results_row_stride,
                                                           char const **error_message);
SZS_B200_API sz_status_t szs_smith_waterman_scores_u64tape(szs_smith_waterman_scores_t engine,
                                                           szs_device_scope_t device,
                                                           sz_sequence_u64tape_t const *queries,
                                                           sz_sequence_u64tape_t const *candidates,
                                                           sz_ssize_t *results, sz_size_t results_row_stride,
                                                           char const **error_message);
SZS_B200_API void szs_smith_waterman_scores_free(szs_smith_waterman_scores_t engine);

/* ---- Fingerprints (reference stringzillas.h:490-598). OUT OF SCOPE: symbols exist so that
 *      bindings resolve; `_init` reports `sz_missing_gpu_k`. ------------------------------------------------------ */

typedef void *szs_fingerprints_t;
SZS_B200_API sz_status_t szs_fingerprints_init(sz_size_t dimensions, sz_size_t alphabet_size,
                                               sz_size_t const *window_widths, sz_size_t window_widths_count,
                                               sz_u64_t seed, sz_memory_allocator_t const *alloc,
                                               sz_capability_t capabilities,
                                               szs_fingerprints_t *engine, char const **error_message);
SZS_B200_API sz_status_t szs_fingerprints_sequence(szs_fingerprints_t engine, szs_device_scope_t device,
                                                   sz_sequence_t const *texts, sz_u32_t *min_hashes,
                                                   sz_size_t min_hashes_stride, sz_u32_t *min_counts,
                                                   sz_size_t min_counts_stride, char const **error_message);
SZS_B200_API sz_status_t szs_fingerprints_u32tape(szs_fingerprints_t engine, szs_device_scope_t device,
                                                  sz_sequence_u32tape_t const *texts, sz_u32_t *min_hashes,
                                                  sz_size_t min_hashes_stride, sz_u32_t *min_counts,
                                                  sz_size_t min_counts_stride, char const **error_message);
SZS_B200_API sz_status_t szs_fingerprints_u64tape(szs_fingerprints_t engine, szs_device_scope_t device,
                                                  sz_sequence_u64tape_t const *texts, sz_u32_t *min_hashes,
                                                  sz_size_t min_hashes_stride, sz_u32_t *min_counts,
                                                  sz_size_t min_counts_stride, char const **error_message);
SZS_B200_API void szs_fingerprints_free(szs_fingerprints_t engine);

/* ---- B200 additions (no reference counterpart) -------------------------------------------------- */

/** Telemetry of the most recent similarity call on this engine (any of the three families). The reference returns
 *  device milliseconds inside `cuda_status_t` to its C++ callers (types.cuh:280-298); the C ABI has no slot for it. */
typedef struct szs_b200_call_stats_t {
    double kernel_milliseconds; /* CUDA-event time of the scoring kernels on the scope's stream */
    double total_milliseconds;  /* staging + kernels + result copy-back, CUDA events */
    sz_u64_t cells;             /* sum over scored pairs of len(query) * len(candidate) */
    sz_u64_t kernel_launches;   /* kernels of this library launched by the call */
    sz_u64_t h2d_bytes;         /* bytes staged host -> device */
    sz_u64_t d2h_bytes;         /* bytes copied device -> host */
    sz_u64_t pairs_u16;         /* (query, candidate) cells scored by the packed int16 DPX tier */
    sz_u64_t pairs_s32;         /* cells scored by the 32-bit DPX tier */
    sz_u64_t pairs_bit_parallel; /* cells scored by the bit-parallel unit-cost Levenshtein kernel */
} szs_b200_call_stats_t;
SZS_B200_API sz_status_t szs_b200_last_call_stats(void const *engine, szs_b200_call_stats_t *stats);

/** The CUDA stream (`cudaStream_t`) every call on this GPU scope enqueues its work on — one stream per scope, like
 *  the reference's `cuda_executor_t` (types.cuh:403-473). Lets a caller bracket calls with its own CUDA events. */
SZS_B200_API sz_status_t szs_b200_device_scope_stream(szs_device_scope_t scope, void **stream, char const **error_message);

/** Unit-cost byte-level Levenshtein normally runs the bit-parallel (Myers/Hyyro) kernel, as every reference backend
 *  does (serial.hpp:2636-2642); `enabled = 0` keeps such engines on the int16 DPX wavefront tier (used by bench.py to
 *  measure that tier on the headline workload). Results are identical either way. */
SZS_B200_API sz_status_t szs_b200_set_bit_parallel(void *engine, int enabled);

/** Diagnostics / unit tests of the host scheduler (no GPU needed): runs the planning steps of a call on string LENGTHS
 *  alone and reports, per candidate, `assignment` = tier (bits 0-7: 0 bit-parallel, 1 packed int16 DP, 2 32-bit DP) |
 *  words-per-lane or strip width (bits 8-15) | lanes (bits 16-23) | family (bits 24-31), its slot `partner` (or
 *  0xFFFFFFFF), the longest-first query `order`, and the DP jobs as (first candidate of the slot, q_begin, q_end)
 *  triples. `candidate_lengths == NULL` plans the symmetric self-similarity of the queries. */
SZS_B200_API sz_status_t szs_b200_debug_schedule(void const *engine, sz_u32_t const *query_lengths, sz_size_t n_queries,
                                                 sz_u32_t const *candidate_lengths, sz_size_t n_candidates,
                                                 int distinct_candidate_bytes, int device_sms, sz_u32_t *assignment,
                                                 sz_u32_t *partner, sz_u32_t *order, sz_u32_t *jobs,
                                                 sz_size_t jobs_capacity, sz_size_t *n_jobs);

/** Score only query rows [row_begin, row_end) of the cross product; `results` still addresses the FULL matrix
 *  (`results[q * stride + c]`), rows outside the range are left untouched. This is the shard primitive of the
 *  multi-GPU partition: each rank owns a contiguous block of query rows balanced by total query length.
 *  `family`: 0 Levenshtein (results are sz_size_t), 1 Needleman-Wunsch, 2 Smith-Waterman (sz_ssize_t).
 *  With `candidates == NULL` (symmetric self-similarity) a row shard scores the cells (query i, candidate j) of its rows
 *  for j < row_end only — the reference's orientation of the lower triangle (serial.hpp:3171-3185) — and the caller
 *  mirrors the upper triangle once every shard is in (the multi-GPU entries below do). */
SZS_B200_API sz_status_t szs_b200_scores_u64tape_rows(int family, void *engine, szs_device_scope_t device,
                                                      sz_sequence_u64tape_t const *queries,
                                                      sz_sequence_u64tape_t const *candidates, sz_size_t row_begin,
                                                      sz_size_t row_end, void *results,
                                                      sz_size_t results_row_stride, char const **error_message);

/* ---- Multi-GPU (the reference has none: `szs_device_scope_init_gpu_device` takes ONE device id,
 *      stringzillas.h:137-138, and no NCCL/MPI/peer-copy symbol exists in its tree) ------------------------------- */

/** Contiguous blocks of query rows of near-equal cost, the partition every multi-GPU entry below uses: a row costs
 *  `len(q) * sum(len(c))`; in a symmetric call (`symmetric != 0`) row i costs `len(i) * sum(len(j), j <= i)`.
 *  Writes `parts + 1` bounds into `bounds`; host-only, no GPU needed. */
SZS_B200_API sz_status_t szs_b200_partition_rows(sz_u64_t const *lengths, sz_size_t count, sz_size_t parts, int symmetric,
                                                 sz_u64_t *bounds);

/** (1) ONE process driving several GPUs: a device group owns one scope (device + stream) and one parked host thread per
 *  GPU; `gpu_devices == NULL` takes every visible device. The reference already allows different scopes on different
 *  host threads (types.cuh:433-444); this is the orchestrator on top of that contract. */
typedef void *szs_b200_device_group_t;
SZS_B200_API sz_status_t szs_b200_device_group_init(sz_size_t const *gpu_devices, sz_size_t count,
                                                    szs_b200_device_group_t *group, char const **error_message);
SZS_B200_API sz_status_t szs_b200_device_group_size(szs_b200_device_group_t group, sz_size_t *count);
SZS_B200_API void szs_b200_device_group_free(szs_b200_device_group_t group);

/** The `*_u64tape` call of `family` (0 Levenshtein, 1 Needleman-Wunsch, 2 Smith-Waterman) sharded by query rows over
 *  the group's GPUs. Every shard reads the inputs straight from the caller's memory (each GPU over its own PCIe link)
 *  and writes its rows straight into the caller's matrix: no collective is involved. `engine` is an ordinary engine
 *  handle; per-device clones of it are kept inside. `candidates == NULL` requests symmetric self-similarity. */
SZS_B200_API sz_status_t szs_b200_scores_u64tape_sharded(int family, void *engine, szs_b200_device_group_t group,
                                                         sz_sequence_u64tape_t const *queries,
                                                         sz_sequence_u64tape_t const *candidates, void *results,
                                                         sz_size_t results_row_stride, char const **error_message);

/** (2) ONE process per GPU (e.g. under `torch.distributed.run`): a communicator wraps an NCCL communicator over the
 *  ranks' GPUs. NCCL is resolved at run time (the copy already in the process, `SZS_B200_NCCL_LIBRARY`, or
 *  libnccl.so.2), so the library loads and scores on one GPU without it. Rank 0 creates the 128-byte identifier and
 *  hands it to the other ranks by any out-of-band means; every rank then calls `szs_b200_comm_init`. */
#define SZS_B200_COMM_ID_BYTES 128
typedef void *szs_b200_comm_t;
SZS_B200_API sz_status_t szs_b200_comm_unique_id(void *unique_id, char const **error_message);
SZS_B200_API sz_status_t szs_b200_comm_init(void const *unique_id, int world_size, int rank, sz_size_t gpu_device,
                                            szs_b200_comm_t *comm, char const **error_message);
SZS_B200_API void szs_b200_comm_free(szs_b200_comm_t comm);

/** Collective call, entered by every rank with its own engine and GPU scope. Rank 0 passes the tapes (host, pinned,
 *  managed or device memory) and the result matrix; the other ranks pass NULL for all three. ONE NCCL broadcast of the
 *  packed inputs (after a 64-byte header), row-sharded scoring, ONE grouped send/recv gather of the int64 score slabs
 *  to rank 0, which copies the whole matrix into `results`. Every (query, candidate) pair is independent, so nothing
 *  crosses GPUs during the DP itself. Arguments rank 0 rejects (null or decreasing offsets, a missing result matrix, a
 *  stride below the candidate count) travel in the header: every rank returns that status instead of waiting in the
 *  broadcast. */
SZS_B200_API sz_status_t szs_b200_comm_scores_u64tape(int family, void *engine, szs_device_scope_t device,
                                                      szs_b200_comm_t comm, sz_sequence_u64tape_t const *queries,
                                                      sz_sequence_u64tape_t const *candidates, void *results,
                                                      sz_size_t results_row_stride, char const **error_message);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* STRINGZILLAS_B200_H_ */
