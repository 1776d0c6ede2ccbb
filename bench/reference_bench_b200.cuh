/**
 *  @file reference_bench_b200.cuh
 *  @brief Adapters that make the engines of libstringzillas_b200.so selectable INSIDE the reference's own CUPS
 *         benchmark (/root/reference/bench/similarities.cu, .cuh:581-639, 848-935): its engines are C++ callables
 *         `engine(queries, candidates, strided_rows<sz_ssize_t> results, cuda_executor_t, gpu_specs_t) -> cuda_status_t`
 *         (cuda.cuh:4419-4431); these classes have that shape and forward to the `szs_*` C ABI of this repo's library.
 *
 *  This file is NOT part of the product and contains no reference code: `oracle/build_ref_bench.sh` copies the
 *  reference's bench sources into oracle/_ref/bench_src/ at build time, splices in the registrations below (a sed
 *  script, see there) and compiles the result. The prototypes come from the REFERENCE's C header
 *  (stringzillas/stringzillas.h) — the adapters are written against the reference's ABI, and the linker resolves them
 *  against this repo's library: the same drop-in a binding maintainer would do. The reference's built-in
 *  `similarities_equality_t` (bench/similarities.cuh:410-423) then compares every tile these engines produce with its
 *  serial engine's, and its "Kernel:" line reads the device time they report.
 */
#pragma once
#include <memory>
#include <span>
#include <stdexcept>
#include <vector>

#include <stringzillas/stringzillas.h> // the REFERENCE's C ABI header

extern "C" {
/* The one B200 addition used here (include/stringzillas_b200.h of this repo): device time of the last call. */
typedef struct szs_b200_call_stats_t {
    double kernel_milliseconds, total_milliseconds;
    sz_u64_t cells, kernel_launches, h2d_bytes, d2h_bytes, pairs_u16, pairs_s32, pairs_bit_parallel;
} szs_b200_call_stats_t;
sz_status_t szs_b200_last_call_stats(void const *engine, szs_b200_call_stats_t *stats);
}

namespace b200_adapters {

namespace szs = ashvardanian::stringzillas;

/** The benchmark's token spans as an `sz_sequence_t` (reference include/stringzilla/types.h:1393-1398). */
template <typename tokens_type_>
struct sequence_of {
    tokens_type_ const *tokens;
    sz_sequence_t sequence;
    explicit sequence_of(tokens_type_ const &t) : tokens(&t) {
        sequence.handle = this, sequence.count = t.size();
        sequence.get_start = [](void const *handle, sz_size_t i) -> sz_cptr_t {
            return (sz_cptr_t)(*static_cast<sequence_of const *>(handle)->tokens)[i].data();
        };
        sequence.get_length = [](void const *handle, sz_size_t i) -> sz_size_t {
            return (sz_size_t)(*static_cast<sequence_of const *>(handle)->tokens)[i].size();
        };
    }
};

struct shared_scope_t {
    szs_device_scope_t scope = nullptr;
    shared_scope_t() {
        char const *message = nullptr;
        if (szs_device_scope_init_gpu_device(0, &scope, &message) != sz_success_k)
            throw std::runtime_error(message ? message : "szs_device_scope_init_gpu_device failed");
    }
    ~shared_scope_t() { szs_device_scope_free(scope); }
};
inline szs_device_scope_t gpu_scope() {
    static shared_scope_t shared;
    return shared.scope;
}

inline szs::cuda_status_t to_status(sz_status_t status, void const *engine) {
    szs::cuda_status_t result;
    result.status = static_cast<szs::status_t>(status);
    szs_b200_call_stats_t stats {};
    if (status == sz_success_k && szs_b200_last_call_stats(engine, &stats) == sz_success_k)
        result.elapsed_milliseconds = static_cast<float>(stats.kernel_milliseconds);
    return result;
}

/** `levenshtein_cuda_t`-shaped engine: uniform substitution costs + linear or affine gaps. */
struct levenshtein_b200_t {
    std::shared_ptr<void> handle;
    levenshtein_b200_t() = default;
    levenshtein_b200_t(szs::uniform_substitution_costs_t subs, szs::linear_gap_costs_t gaps) { init(subs, gaps.open_or_extend, gaps.open_or_extend); }
    levenshtein_b200_t(szs::uniform_substitution_costs_t subs, szs::affine_gap_costs_t gaps) { init(subs, gaps.open, gaps.extend); }
    void init(szs::uniform_substitution_costs_t subs, sz_error_cost_t open, sz_error_cost_t extend) {
        szs_levenshtein_distances_t raw = nullptr;
        char const *message = nullptr;
        if (szs_levenshtein_distances_init(subs.match, subs.mismatch, open, extend, nullptr, sz_caps_cuda_k, &raw, &message) != sz_success_k)
            throw std::runtime_error(message ? message : "szs_levenshtein_distances_init failed");
        handle.reset(raw, [](void *engine) { szs_levenshtein_distances_free(engine); });
    }
    template <typename queries_type_, typename candidates_type_, typename... ignored_types_>
    szs::cuda_status_t operator()(queries_type_ const &queries, candidates_type_ const &candidates,
                                  szs::strided_rows<sz_ssize_t> results, ignored_types_ const &...) const noexcept {
        sequence_of<queries_type_> q(queries);
        sequence_of<candidates_type_> c(candidates);
        char const *message = nullptr;
        sz_status_t const status = szs_levenshtein_distances(handle.get(), gpu_scope(), &q.sequence, &c.sequence,
                                                             (sz_size_t *)results.data, results.row_stride, &message);
        return to_status(status, handle.get());
    }
};
using affine_levenshtein_b200_t = levenshtein_b200_t;

/** `needleman_wunsch_cuda_t` / `smith_waterman_cuda_t`-shaped engines: 32-class substitution table + gaps. */
template <bool local_>
struct weighted_b200_t {
    std::shared_ptr<void> handle;
    weighted_b200_t() = default;
    weighted_b200_t(szs::error_costs_32x32_t const &subs, szs::linear_gap_costs_t gaps) { init(subs, gaps.open_or_extend, gaps.open_or_extend); }
    weighted_b200_t(szs::error_costs_32x32_t const &subs, szs::affine_gap_costs_t gaps) { init(subs, gaps.open, gaps.extend); }
    void init(szs::error_costs_32x32_t const &subs, sz_error_cost_t open, sz_error_cost_t extend) {
        void *raw = nullptr;
        char const *message = nullptr;
        sz_status_t const status =
            local_ ? szs_smith_waterman_scores_init(subs.byte_to_class, &subs.class_substitution_costs[0][0], open, extend, nullptr, sz_caps_cuda_k, &raw, &message)
                   : szs_needleman_wunsch_scores_init(subs.byte_to_class, &subs.class_substitution_costs[0][0], open, extend, nullptr, sz_caps_cuda_k, &raw, &message);
        if (status != sz_success_k) throw std::runtime_error(message ? message : "szs_*_scores_init failed");
        handle.reset(raw, [](void *engine) { local_ ? szs_smith_waterman_scores_free(engine) : szs_needleman_wunsch_scores_free(engine); });
    }
    template <typename queries_type_, typename candidates_type_, typename... ignored_types_>
    szs::cuda_status_t operator()(queries_type_ const &queries, candidates_type_ const &candidates,
                                  szs::strided_rows<sz_ssize_t> results, ignored_types_ const &...) const noexcept {
        sequence_of<queries_type_> q(queries);
        sequence_of<candidates_type_> c(candidates);
        char const *message = nullptr;
        sz_status_t const status =
            local_ ? szs_smith_waterman_scores(handle.get(), gpu_scope(), &q.sequence, &c.sequence, results.data, results.row_stride, &message)
                   : szs_needleman_wunsch_scores(handle.get(), gpu_scope(), &q.sequence, &c.sequence, results.data, results.row_stride, &message);
        return to_status(status, handle.get());
    }
};
using needleman_wunsch_b200_t = weighted_b200_t<false>;
using affine_needleman_wunsch_b200_t = weighted_b200_t<false>;
using smith_waterman_b200_t = weighted_b200_t<true>;
using affine_smith_waterman_b200_t = weighted_b200_t<true>;

} // namespace b200_adapters
