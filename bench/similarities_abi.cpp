/**
 *  CUPS benchmark for any library exporting the `szs_*` similarity ABI, in the terms of the reference's own
 *  `bench/similarities.cu`: same environment variables, the same tile shapes, the same cost schemes,
 *  the same cell accounting, the same "accelerated must equal the serial baseline" stress check, the same log lines
 *  (`> Throughput`, `> Efficiency … GOps/s`, `> Kernel: … GCUPS`).
 *
 *  The reference's bench instantiates its C++ engine templates; it cannot see an engine that lives behind the C ABI.
 *  This program sits one level higher: both arms are shared libraries opened at run time,
 *
 *      SZS_ACCELERATED_LIB   default  <repo>/stringzilla_b200/lib/libstringzillas_b200.so   (GPU scope, device 0)
 *      SZS_BASELINE_LIB      default  <repo>/oracle/_ref/libstringzillas_ref.so             (CPU scope, all cores)
 *
 *  so the same binary also times the reference's CUDA build against its CPU build, or two builds of this repo.
 *
 *  Environment (reference bench/similarities.cu:16-28, bench/similarities.cuh:181-244):
 *      STRINGWARS_DATASET         path of a text file; without it a synthetic corpus is generated (see below)
 *      STRINGWARS_TOKENS          "lines" (default) | "words" | "file" | N  (N-byte slices)
 *      STRINGWARS_SEED            default 42
 *      STRINGWARS_DURATION        seconds per benchmark, default 10 (fractions allowed)
 *      STRINGWARS_STRESS          1 = compare accelerated results with the baseline's on identical tiles
 *      STRINGWARS_STRESS_DURATION seconds per stress run, default 10
 *      STRINGWARS_STRESS_LIMIT    tolerated mismatching tiles, default 1
 *      STRINGWARS_FILTER          regular expression over benchmark names
 *      STRINGWARS_BATCH           comma-separated query counts overriding the shape sweep
 *      STRINGWARS_BATCH_PER_CORE  pairs per core (GPU SM or CPU thread), default 256
 *      SZS_CPU_CORES              threads of a CPU device scope, default 0 = all (the reference's pool busy-waits)
 *      STRINGWARS_SYNTHETIC_LENGTH  mean token length of the synthetic corpus, default 1024 (lengths U[0.94L, 1.06L])
 *
 *  Synthetic corpus: `ACGT` tokens for the Levenshtein benchmarks, 20-amino-acid tokens for NW/SW, the methodology of the
 *  reference's published table (include/stringzillas/similarities/README.md:12).
 */
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <regex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <string_view>
#include <thread>
#include <type_traits>
#include <vector>

#include "stringzillas_b200.h"

namespace {

using clock_type = std::chrono::steady_clock;

double seconds_since(clock_type::time_point start) {
    return std::chrono::duration<double>(clock_type::now() - start).count();
}

std::string env_or(char const *name, std::string fallback) {
    char const *value = std::getenv(name);
    return value && *value ? std::string(value) : fallback;
}

/* ------------------------------------------------------------------ the ABI, resolved at run time ------------ */

struct Library {
    void *handle = nullptr;
    std::string path;

    template <typename signature_> signature_ *symbol(char const *name, bool required = true) const {
        void *address = dlsym(handle, name);
        if (!address && required) throw std::runtime_error(path + " does not export " + name);
        return reinterpret_cast<signature_ *>(address);
    }

    // Every signature is taken from the header's own declaration, so an argument in the wrong place cannot compile.
    using init_levenshtein_t = std::remove_pointer_t<decltype(&szs_levenshtein_distances_init)>;
    using init_weighted_t = std::remove_pointer_t<decltype(&szs_needleman_wunsch_scores_init)>;
    using call_distances_t = std::remove_pointer_t<decltype(&szs_levenshtein_distances_u64tape)>;
    using call_scores_t = std::remove_pointer_t<decltype(&szs_needleman_wunsch_scores_u64tape)>;
    using free_t = std::remove_pointer_t<decltype(&szs_levenshtein_distances_free)>;
    static_assert(std::is_same_v<init_levenshtein_t, std::remove_pointer_t<decltype(&szs_levenshtein_distances_utf8_init)>>);
    static_assert(std::is_same_v<init_weighted_t, std::remove_pointer_t<decltype(&szs_smith_waterman_scores_init)>>);
    static_assert(std::is_same_v<call_distances_t, std::remove_pointer_t<decltype(&szs_levenshtein_distances_utf8_u64tape)>>);
    static_assert(std::is_same_v<call_scores_t, std::remove_pointer_t<decltype(&szs_smith_waterman_scores_u64tape)>>);

    decltype(&szs_capabilities) capabilities = nullptr;
    decltype(&szs_device_scope_init_cpu_cores) scope_cpu = nullptr;
    decltype(&szs_device_scope_init_gpu_device) scope_gpu = nullptr;
    decltype(&szs_device_scope_free) scope_free = nullptr;
    init_levenshtein_t *levenshtein_init = nullptr, *levenshtein_utf8_init = nullptr;
    init_weighted_t *needleman_wunsch_init = nullptr, *smith_waterman_init = nullptr;
    call_distances_t *levenshtein = nullptr, *levenshtein_utf8 = nullptr;
    call_scores_t *needleman_wunsch = nullptr, *smith_waterman = nullptr;
    free_t *levenshtein_free = nullptr, *levenshtein_utf8_free = nullptr, *needleman_wunsch_free = nullptr,
           *smith_waterman_free = nullptr;
    decltype(&szs_b200_last_call_stats) last_call_stats = nullptr; // only this repo's library exports it

    szs_device_scope_t scope = nullptr;
    bool on_gpu = false;

    explicit Library(std::string const &library_path) : path(library_path) {
        handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!handle) throw std::runtime_error(std::string("dlopen failed: ") + dlerror());
        capabilities = symbol<std::remove_pointer_t<decltype(capabilities)>>("szs_capabilities");
        scope_cpu = symbol<std::remove_pointer_t<decltype(scope_cpu)>>("szs_device_scope_init_cpu_cores");
        scope_gpu = symbol<std::remove_pointer_t<decltype(scope_gpu)>>("szs_device_scope_init_gpu_device");
        scope_free = symbol<std::remove_pointer_t<decltype(scope_free)>>("szs_device_scope_free");
        levenshtein_init = symbol<init_levenshtein_t>("szs_levenshtein_distances_init");
        levenshtein_utf8_init = symbol<init_levenshtein_t>("szs_levenshtein_distances_utf8_init");
        needleman_wunsch_init = symbol<init_weighted_t>("szs_needleman_wunsch_scores_init");
        smith_waterman_init = symbol<init_weighted_t>("szs_smith_waterman_scores_init");
        levenshtein = symbol<call_distances_t>("szs_levenshtein_distances_u64tape");
        levenshtein_utf8 = symbol<call_distances_t>("szs_levenshtein_distances_utf8_u64tape");
        needleman_wunsch = symbol<call_scores_t>("szs_needleman_wunsch_scores_u64tape");
        smith_waterman = symbol<call_scores_t>("szs_smith_waterman_scores_u64tape");
        levenshtein_free = symbol<free_t>("szs_levenshtein_distances_free");
        levenshtein_utf8_free = symbol<free_t>("szs_levenshtein_distances_utf8_free");
        needleman_wunsch_free = symbol<free_t>("szs_needleman_wunsch_scores_free");
        smith_waterman_free = symbol<free_t>("szs_smith_waterman_scores_free");
        last_call_stats = symbol<std::remove_pointer_t<decltype(last_call_stats)>>("szs_b200_last_call_stats", false);
    }

    /** GPU scope when the library reports CUDA capabilities and a device answers, else all CPU cores. */
    void open_scope(std::size_t cpu_cores) {
        char const *why = nullptr;
        if ((capabilities() & sz_cap_cuda_k) && scope_gpu(0, &scope, &why) == sz_success_k) {
            on_gpu = true;
            return;
        }
        if (scope_cpu(cpu_cores, &scope, &why) != sz_success_k)
            throw std::runtime_error(path + ": no usable device scope: " + (why ? why : "?"));
    }

    /** Engines are built for the tier the scope will run: the library's CUDA bits on a GPU (the reference refuses a
     *  mask mixing CPU tiers into a GPU engine), its CPU bits otherwise. */
    sz_capability_t engine_capabilities() const {
        unsigned const all = static_cast<unsigned>(capabilities());
        unsigned const cuda_bits = static_cast<unsigned>(sz_cap_cuda_k) | sz_cap_kepler_k | sz_cap_hopper_k;
        return static_cast<sz_capability_t>(on_gpu ? (all & cuda_bits) : (all & ~cuda_bits));
    }

    ~Library() {
        if (scope) scope_free(scope);
        // The handle stays open: CUDA runtimes linked into the libraries dislike being unloaded at exit.
    }
};

/** Streaming-multiprocessor count of device 0 through the driver API, 0 when there is no driver on this host. */
std::size_t streaming_multiprocessors() {
    void *driver = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!driver) return 0;
    auto init = reinterpret_cast<int (*)(unsigned)>(dlsym(driver, "cuInit"));
    auto get = reinterpret_cast<int (*)(int *, int)>(dlsym(driver, "cuDeviceGet"));
    auto attribute = reinterpret_cast<int (*)(int *, int, int)>(dlsym(driver, "cuDeviceGetAttribute"));
    int device = 0, count = 0;
    int const multiprocessor_count_attribute = 16; // CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT
    if (!init || !get || !attribute || init(0) != 0 || get(&device, 0) != 0 ||
        attribute(&count, multiprocessor_count_attribute, device) != 0)
        return 0;
    return static_cast<std::size_t>(count);
}

/* ------------------------------------------------------------------ corpus and tiles --------------------------- */

struct Corpus {
    std::string blob;
    std::vector<std::string_view> tokens;
    std::string description;

    double mean_length() const {
        double total = 0;
        for (auto token : tokens) total += static_cast<double>(token.size());
        return tokens.empty() ? 0 : total / static_cast<double>(tokens.size());
    }
};

Corpus tokenize(std::string blob, std::string const &mode, std::uint64_t seed, std::string description) {
    Corpus corpus;
    corpus.blob = std::move(blob);
    corpus.description = std::move(description);
    std::string_view const text = corpus.blob;
    auto split = [&](auto is_separator) {
        std::size_t begin = 0;
        for (std::size_t i = 0; i <= text.size(); ++i)
            if (i == text.size() || is_separator(text[i])) {
                if (i > begin) corpus.tokens.push_back(text.substr(begin, i - begin));
                begin = i + 1;
            }
    };
    if (mode == "file") corpus.tokens.push_back(text);
    else if (mode == "lines") split([](char c) { return c == '\n' || c == '\r'; });
    else if (mode == "words") split([](char c) { return c == ' ' || c == '\n' || c == '\r' || c == '\t'; });
    else {
        std::size_t const width = std::stoull(mode);
        if (!width) throw std::runtime_error("STRINGWARS_TOKENS must be lines, words, file or a positive integer");
        for (std::size_t begin = 0; begin + width <= text.size(); begin += width)
            corpus.tokens.push_back(text.substr(begin, width));
    }
    std::mt19937_64 generator(seed);
    std::shuffle(corpus.tokens.begin(), corpus.tokens.end(), generator);
    return corpus;
}

Corpus synthesize(std::string_view alphabet, std::size_t mean_length, std::size_t count, std::uint64_t seed) {
    std::mt19937_64 generator(seed);
    std::size_t const low = mean_length - mean_length / 16, high = mean_length + mean_length / 16;
    std::uniform_int_distribution<std::size_t> length(low, high), letter(0, alphabet.size() - 1);
    std::string blob;
    blob.reserve(count * (high + 1));
    for (std::size_t i = 0; i < count; ++i) {
        for (std::size_t n = length(generator); n; --n) blob.push_back(alphabet[letter(generator)]);
        blob.push_back('\n');
    }
    std::ostringstream description;
    description << "synthetic, alphabet `" << alphabet << "`, lengths U[" << low << "," << high << "]";
    return tokenize(std::move(blob), "lines", seed, description.str());
}

struct Shape {
    std::size_t queries = 1, candidates = 1;
    bool all_pairs = true;

    std::size_t scored_pairs() const { return all_pairs ? queries * candidates : std::min(queries, candidates); }
    std::string label() const {
        return all_pairs ? "q" + std::to_string(queries) + "xc" + std::to_string(candidates)
                         : "pairs" + std::to_string(scored_pairs());
    }
};

/** The sweep of bench/similarities.cuh:191-244: a pair budget of `per_core * cores` split as a retrieval corner, a
 *  candidate-heavy skew, the balanced square and a query-heavy skew, each beside a diagonal of independent 1x1 calls. */
std::vector<Shape> make_shapes(Corpus const &corpus, std::size_t cores) {
    std::size_t const per_core = std::max<std::size_t>(1, std::stoull(env_or("STRINGWARS_BATCH_PER_CORE", "256")));
    std::size_t const budget = per_core * std::max<std::size_t>(1, cores);
    std::size_t side = 1;
    while ((side + 1) * (side + 1) <= budget) ++side;
    std::size_t const tokens = corpus.tokens.size();
    std::size_t const largest_block = tokens > 1 ? tokens / 2 : 1;
    auto clamp = [&](std::size_t value) { return std::max<std::size_t>(1, std::min(value, largest_block)); };
    auto cross = [&](std::size_t wanted) {
        Shape shape;
        shape.queries = clamp(wanted);
        shape.candidates = clamp(std::max<std::size_t>(1, budget / shape.queries));
        if (tokens > shape.queries + 1) shape.candidates = std::min(shape.candidates, tokens - shape.queries - 1);
        shape.candidates = std::max<std::size_t>(1, shape.candidates);
        return shape;
    };
    auto diagonal = [&](std::size_t wanted) {
        std::size_t const half = tokens > 2 ? (tokens - 1) / 2 : 1;
        Shape shape;
        shape.queries = shape.candidates = std::max<std::size_t>(1, std::min(clamp(wanted), half));
        shape.all_pairs = false;
        return shape;
    };
    std::vector<std::size_t> wanted;
    std::string const override_list = env_or("STRINGWARS_BATCH", "");
    if (!override_list.empty()) {
        std::istringstream stream(override_list);
        for (std::string item; std::getline(stream, item, ',');) wanted.push_back(std::stoull(item));
    }
    else wanted = {1, std::max<std::size_t>(1, side / 8), side, side * 8};
    std::vector<Shape> shapes;
    for (std::size_t queries : wanted) shapes.push_back(cross(queries)), shapes.push_back(diagonal(queries));
    return shapes;
}

/** A packed 64-bit tape over a contiguous block of tokens: what `szs_*_u64tape` consumes. */
struct Tape {
    std::vector<char> data;
    std::vector<sz_u64_t> offsets;

    void assign(std::string_view const *tokens, std::size_t count) {
        data.clear(), offsets.assign(1, 0);
        for (std::size_t i = 0; i < count; ++i) {
            data.insert(data.end(), tokens[i].begin(), tokens[i].end());
            offsets.push_back(data.size());
        }
        if (data.empty()) data.push_back(0); // keep `data()` dereferenceable for all-empty blocks
    }
    sz_sequence_u64tape_t view(std::size_t first = 0, std::size_t count = ~std::size_t(0)) const {
        sz_sequence_u64tape_t tape;
        tape.data = data.data();
        tape.offsets = offsets.data() + first;
        tape.count = std::min(count, offsets.size() - 1 - first);
        return tape;
    }
};

/* ------------------------------------------------------------------ one benchmarkable engine ------------------- */

struct Outcome {
    std::string name;
    bool skipped = true;
    double seconds = 0, cells = 0, bytes = 0, kernel_milliseconds = 0, kernel_cells = 0;
    std::size_t calls = 0, pairs = 0, stress_calls = 0, errors = 0;
    std::int64_t last_tile_min = 0, last_tile_max = 0; // extremes of the last scored tile: a sanity line in the log
};

struct Runner {
    Library &library;
    Corpus const &corpus;
    Shape shape;
    Library::call_distances_t *call_distances = nullptr; // exactly one of the two is set
    Library::call_scores_t *call_scores = nullptr;
    Library::free_t *release = nullptr;
    void *engine = nullptr;
    std::uint64_t seed = 42;
    std::vector<std::int64_t> results; // `sz_size_t` distances and `sz_ssize_t` scores share the 8-byte cell
    Tape queries, candidates;

    Runner(Library &lib, Corpus const &text, Shape tile, std::uint64_t random_seed)
        : library(lib), corpus(text), shape(tile), seed(random_seed) {
        if (corpus.tokens.size() <= shape.queries + shape.candidates)
            throw std::runtime_error("the tile does not fit into the corpus");
        results.assign(shape.all_pairs ? shape.queries * shape.candidates : shape.scored_pairs(), 0);
    }
    Runner(Runner const &) = delete;
    ~Runner() {
        if (engine) release(engine);
    }

    std::size_t block_start(std::size_t block, std::uint64_t salt) const {
        std::mt19937_64 generator(seed ^ salt);
        return std::uniform_int_distribution<std::size_t>(0, corpus.tokens.size() - block)(generator);
    }

    /** Scores tile number `index`: two blocks of the corpus at offsets that depend only on (seed, index), so both arms
     *  of a comparison see the same strings (bench/similarities.cuh:318-342 has the same property). */
    void run(std::size_t index, Outcome &outcome) {
        std::uint64_t const salt = static_cast<std::uint64_t>(index) * 0x9E3779B97F4A7C15ull;
        std::size_t const first_query = block_start(shape.queries, salt ^ 0x1111111111111111ull);
        std::size_t first_candidate = block_start(shape.candidates, salt ^ 0x2222222222222222ull);
        bool const overlap = first_candidate + shape.candidates > first_query && first_query + shape.queries > first_candidate;
        if (overlap) {
            std::size_t const after = first_query + shape.queries;
            first_candidate = after + shape.candidates <= corpus.tokens.size() ? after
                              : first_query >= shape.candidates                ? first_query - shape.candidates
                                                                               : 0;
        }
        std::string_view const *query_tokens = corpus.tokens.data() + first_query;
        std::string_view const *candidate_tokens = corpus.tokens.data() + first_candidate;
        queries.assign(query_tokens, shape.queries);
        candidates.assign(candidate_tokens, shape.candidates);

        auto invoke = [&](sz_sequence_u64tape_t q, sz_sequence_u64tape_t c, std::int64_t *out, std::size_t stride) {
            char const *why = nullptr;
            sz_status_t const status =
                call_distances ? call_distances(engine, library.scope, &q, &c, reinterpret_cast<sz_size_t *>(out), stride, &why)
                               : call_scores(engine, library.scope, &q, &c, reinterpret_cast<sz_ssize_t *>(out), stride, &why);
            if (status != sz_success_k)
                throw std::runtime_error(outcome.name + " failed with status " + std::to_string(status) + ": " +
                                         (why ? why : ""));
            if (library.last_call_stats) {
                szs_b200_call_stats_t stats;
                if (library.last_call_stats(engine, &stats) == sz_success_k) {
                    outcome.kernel_milliseconds += stats.kernel_milliseconds;
                    outcome.kernel_cells += static_cast<double>(stats.cells);
                }
            }
        };
        if (shape.all_pairs) {
            invoke(queries.view(), candidates.view(), results.data(), shape.candidates);
            double query_bytes = 0, candidate_bytes = 0;
            for (std::size_t i = 0; i < shape.queries; ++i) query_bytes += static_cast<double>(query_tokens[i].size());
            for (std::size_t j = 0; j < shape.candidates; ++j)
                candidate_bytes += static_cast<double>(candidate_tokens[j].size());
            outcome.cells += query_bytes * candidate_bytes;
            outcome.bytes += query_bytes * static_cast<double>(shape.candidates) +
                             candidate_bytes * static_cast<double>(shape.queries);
        }
        else
            for (std::size_t pair = 0; pair < shape.scored_pairs(); ++pair) {
                invoke(queries.view(pair, 1), candidates.view(pair, 1), results.data() + pair, 1);
                outcome.cells += static_cast<double>(query_tokens[pair].size()) *
                                 static_cast<double>(candidate_tokens[pair].size());
                outcome.bytes += static_cast<double>(query_tokens[pair].size() + candidate_tokens[pair].size());
            }
        outcome.pairs += shape.scored_pairs();
        outcome.calls += 1;
        auto const extremes = std::minmax_element(results.begin(), results.end());
        outcome.last_tile_min = *extremes.first, outcome.last_tile_max = *extremes.second;
    }
};

struct Settings {
    double seconds = 10, stress_seconds = 10;
    bool stress = false;
    std::size_t stress_limit = 1;
    std::uint64_t seed = 42;
    std::string filter;

    bool allows(std::string const &name) const {
        return filter.empty() || std::regex_search(name, std::regex(filter));
    }
};

void print_outcome(Outcome const &outcome, Outcome const *against) {
    std::printf("\nBenchmarking `%s`:\n", outcome.name.c_str());
    if (outcome.stress_calls) std::printf("> Stress: %zu tiles compared with the baseline\n", outcome.stress_calls);
    if (outcome.errors) std::printf("> Errors: %zu in %zu calls\n", outcome.errors, outcome.stress_calls);
    double const per_call = outcome.seconds / static_cast<double>(std::max<std::size_t>(1, outcome.calls));
    std::printf("> Throughput: %.2f MB/s @ %.3f ms/call, %zu calls, %zu pairs\n", outcome.bytes / outcome.seconds / 1e6,
                per_call * 1e3, outcome.calls, outcome.pairs);
    std::printf("> Efficiency: %.2f GOps/s (cell updates per second, wall clock through the C ABI)\n",
                outcome.cells / outcome.seconds / 1e9);
    std::printf("> Results of the last tile: min %lld, max %lld\n", static_cast<long long>(outcome.last_tile_min),
                static_cast<long long>(outcome.last_tile_max));
    if (outcome.kernel_milliseconds > 0)
        std::printf("> Kernel: %.2f GCUPS @ %.3f ms device-measured (excludes host materialization)\n",
                    outcome.kernel_cells / (outcome.kernel_milliseconds * 1e6), outcome.kernel_milliseconds);
    if (against && !against->skipped && against->seconds > 0) {
        double const ratio = (outcome.cells / outcome.seconds) / (against->cells / against->seconds);
        std::printf("> %s%.1f x against `%s`\n", ratio >= 1 ? "+ " : "- ", ratio >= 1 ? ratio : 1 / ratio,
                    against->name.c_str());
    }
    std::fflush(stdout);
}

/** Times `runner` for `settings.seconds`; when a `baseline` runner is given and stress-testing is on, first compares
 *  the two result buffers on identical tiles (bench/shared.hpp:846-865 does the same for the reference's backends). */
Outcome benchmark(Settings const &settings, std::string const &name, Runner &runner, Runner *baseline) {
    Outcome outcome;
    outcome.name = name;
    if (!settings.allows(name)) return outcome;
    outcome.skipped = false;

    if (baseline && settings.stress) {
        Outcome scratch_accelerated, scratch_baseline;
        scratch_accelerated.name = name, scratch_baseline.name = name + " (baseline)";
        auto const start = clock_type::now();
        for (std::size_t tile = 0; tile == 0 || seconds_since(start) < settings.stress_seconds; ++tile) {
            runner.run(tile, scratch_accelerated);
            baseline->run(tile, scratch_baseline);
            ++outcome.stress_calls;
            if (runner.results == baseline->results) continue;
            ++outcome.errors;
            for (std::size_t cell = 0; cell < runner.results.size(); ++cell)
                if (runner.results[cell] != baseline->results[cell]) {
                    std::printf("Mismatch in `%s`, tile %zu, cell %zu: %lld != %lld\n", name.c_str(), tile, cell,
                                static_cast<long long>(runner.results[cell]),
                                static_cast<long long>(baseline->results[cell]));
                    break;
                }
            if (outcome.errors > settings.stress_limit)
                throw std::runtime_error("too many mismatching tiles in " + name);
        }
    }

    Outcome warmup;
    warmup.name = name;
    runner.run(0, warmup);
    auto const start = clock_type::now();
    for (std::size_t tile = 0; tile == 0 || seconds_since(start) < settings.seconds; ++tile) runner.run(tile, outcome);
    outcome.seconds = seconds_since(start);
    return outcome;
}

/* ------------------------------------------------------------------ the two benchmark families ----------------- */

struct CostScheme {
    sz_error_cost_t match, mismatch, linear_gap, affine_open, affine_extend;
    char const *tag;
};
// bench/similarities.cuh:464-467: unit costs take the bit-parallel path, the "weird" ones force the general scorers.
CostScheme const kCostSchemes[] = {{0, 1, 1, 1, 1, "unit"}, {1, 3, 3, 4, 2, "weird"}};

std::string backend_tag(Library const &library, bool is_baseline) {
    if (is_baseline) return library.on_gpu ? "baseline_gpu" : "serial";
    return library.last_call_stats ? "b200" : (library.on_gpu ? "cuda" : "cpu");
}

void bench_levenshtein(Settings const &settings, Corpus const &corpus, std::vector<Shape> const &shapes,
                       Library &accelerated, Library *baseline) {
    struct Variant {
        char const *prefix;
        bool affine, utf8;
    };
    Variant const variants[] = {{"levenshtein", false, false}, {"affine_levenshtein", true, false},
                                {"levenshtein_utf8", false, true}};
    for (CostScheme const &scheme : kCostSchemes)
        for (Shape const &shape : shapes)
            for (Variant const &variant : variants) {
                auto prepare = [&](Library &library, Runner &runner) {
                    char const *why = nullptr;
                    auto init = variant.utf8 ? library.levenshtein_utf8_init : library.levenshtein_init;
                    sz_error_cost_t const open = variant.affine ? scheme.affine_open : scheme.linear_gap;
                    sz_error_cost_t const extend = variant.affine ? scheme.affine_extend : scheme.linear_gap;
                    sz_status_t const status = init(scheme.match, scheme.mismatch, open, extend, nullptr,
                                                    library.engine_capabilities(), &runner.engine, &why);
                    if (status != sz_success_k)
                        throw std::runtime_error(std::string("engine construction failed: ") + (why ? why : "?"));
                    runner.call_distances = variant.utf8 ? library.levenshtein_utf8 : library.levenshtein;
                    runner.release = variant.utf8 ? library.levenshtein_utf8_free : library.levenshtein_free;
                };
                auto name_of = [&](Library const &library, bool is_baseline) {
                    return std::string(variant.prefix) + "_" + backend_tag(library, is_baseline) + "_" + scheme.tag + ":" +
                           shape.label();
                };
                std::string const accelerated_name = name_of(accelerated, false);
                std::string const baseline_name = baseline ? name_of(*baseline, true) : std::string();
                if (!settings.allows(accelerated_name) && !(baseline && settings.allows(baseline_name))) continue;

                Outcome baseline_outcome;
                Runner baseline_runner(baseline ? *baseline : accelerated, corpus, shape, settings.seed);
                if (baseline) {
                    prepare(*baseline, baseline_runner);
                    baseline_outcome = benchmark(settings, baseline_name, baseline_runner, nullptr);
                    if (!baseline_outcome.skipped) print_outcome(baseline_outcome, nullptr);
                }
                Runner runner(accelerated, corpus, shape, settings.seed);
                prepare(accelerated, runner);
                Outcome const outcome = benchmark(settings, accelerated_name, runner, baseline ? &baseline_runner : nullptr);
                if (!outcome.skipped) print_outcome(outcome, baseline ? &baseline_outcome : nullptr);
            }
}

/** BLOSUM62 in the engines' compact layout (256-entry byte -> class map, 32x32 class table), generated from the Python
 *  host layer's `scoring.blosum62()`, the restatement of the reference's `error_costs_32x32_t::blosum62()`. */
#include "blosum62.inc"

void bench_needleman_wunsch_smith_waterman(Settings const &settings, Corpus const &corpus,
                                           std::vector<Shape> const &shapes, Library &accelerated, Library *baseline) {
    sz_error_cost_t const *table = kBlosum62Costs;
    sz_u8_t const *byte_to_class = kBlosum62ByteToClass;
    struct Variant {
        char const *prefix;
        bool local, affine;
    };
    // bench/similarities.cuh:647-649: linear gap -4, affine open -4 / extend -1
    Variant const variants[] = {{"needleman_wunsch", false, false}, {"smith_waterman", true, false},
                                {"affine_needleman_wunsch", false, true}, {"affine_smith_waterman", true, true}};
    for (Shape const &shape : shapes)
        for (Variant const &variant : variants) {
            auto prepare = [&](Library &library, Runner &runner) {
                char const *why = nullptr;
                auto init = variant.local ? library.smith_waterman_init : library.needleman_wunsch_init;
                sz_status_t const status = init(byte_to_class, table, -4, variant.affine ? -1 : -4, nullptr,
                                                library.engine_capabilities(), &runner.engine, &why);
                if (status != sz_success_k)
                    throw std::runtime_error(std::string("engine construction failed: ") + (why ? why : "?"));
                runner.call_scores = variant.local ? library.smith_waterman : library.needleman_wunsch;
                runner.release = variant.local ? library.smith_waterman_free : library.needleman_wunsch_free;
            };
            auto name_of = [&](Library const &library, bool is_baseline) {
                return std::string(variant.prefix) + "_" + backend_tag(library, is_baseline) + ":" + shape.label();
            };
            std::string const accelerated_name = name_of(accelerated, false);
            std::string const baseline_name = baseline ? name_of(*baseline, true) : std::string();
            if (!settings.allows(accelerated_name) && !(baseline && settings.allows(baseline_name))) continue;

            Outcome baseline_outcome;
            Runner baseline_runner(baseline ? *baseline : accelerated, corpus, shape, settings.seed);
            if (baseline) {
                prepare(*baseline, baseline_runner);
                baseline_outcome = benchmark(settings, baseline_name, baseline_runner, nullptr);
                if (!baseline_outcome.skipped) print_outcome(baseline_outcome, nullptr);
            }
            Runner runner(accelerated, corpus, shape, settings.seed);
            prepare(accelerated, runner);
            Outcome const outcome = benchmark(settings, accelerated_name, runner, baseline ? &baseline_runner : nullptr);
            if (!outcome.skipped) print_outcome(outcome, baseline ? &baseline_outcome : nullptr);
        }
}

std::string repository_root(char const *program) {
    std::string path = program;
    std::size_t const slash = path.find_last_of('/');
    std::string const directory = slash == std::string::npos ? "." : path.substr(0, slash);
    return directory + "/.."; // the binary lives in <repo>/bench/
}

} // namespace

int main(int, char const **argv) {
    try {
        std::string const root = repository_root(argv[0]);
        Settings settings;
        settings.seconds = std::stod(env_or("STRINGWARS_DURATION", "10"));
        settings.stress = env_or("STRINGWARS_STRESS", "0") == "1";
        settings.stress_seconds = std::stod(env_or("STRINGWARS_STRESS_DURATION", "10"));
        settings.stress_limit = std::stoull(env_or("STRINGWARS_STRESS_LIMIT", "1"));
        settings.seed = std::stoull(env_or("STRINGWARS_SEED", "42"));
        settings.filter = env_or("STRINGWARS_FILTER", "");

        Library accelerated(env_or("SZS_ACCELERATED_LIB", root + "/stringzilla_b200/lib/libstringzillas_b200.so"));
        // The reference's CPU pool busy-waits between calls: on a host with a CPU quota below its core count, cap it.
        std::size_t const cpu_cores = std::stoull(env_or("SZS_CPU_CORES", "0")); // 0 = every core the library sees
        accelerated.open_scope(cpu_cores);
        std::string const baseline_path = env_or("SZS_BASELINE_LIB", root + "/oracle/_ref/libstringzillas_ref.so");
        Library *baseline = nullptr;
        if (baseline_path != "none" && std::ifstream(baseline_path).good()) {
            baseline = new Library(baseline_path);
            baseline->open_scope(cpu_cores);
        }

        std::size_t cores = accelerated.on_gpu ? streaming_multiprocessors() : 0;
        if (!cores) cores = std::max(1u, std::thread::hardware_concurrency());
        std::printf("Accelerated: %s on %s (capabilities 0x%x)\n", accelerated.path.c_str(),
                    accelerated.on_gpu ? "GPU 0" : "CPU cores", static_cast<unsigned>(accelerated.engine_capabilities()));
        if (baseline)
            std::printf("Baseline:    %s on %s (capabilities 0x%x)\n", baseline->path.c_str(),
                        baseline->on_gpu ? "GPU 0" : "CPU cores", static_cast<unsigned>(baseline->engine_capabilities()));
        else std::printf("Baseline:    none (no stress comparison, no relative figures)\n");
        std::printf("Cores for batch sizing: %zu (%s)\n", cores,
                    accelerated.on_gpu ? "streaming multiprocessors" : "hardware threads");

        std::string const dataset = env_or("STRINGWARS_DATASET", "");
        std::string const tokens = env_or("STRINGWARS_TOKENS", "lines");
        std::size_t const synthetic_length = std::stoull(env_or("STRINGWARS_SYNTHETIC_LENGTH", "1024"));
        auto report = [](char const *purpose, Corpus const &corpus) {
            std::printf("Corpus for %s: %s; %zu tokens, mean length %.2f bytes\n", purpose, corpus.description.c_str(),
                        corpus.tokens.size(), corpus.mean_length());
        };
        auto load = [&](std::string_view alphabet) {
            if (dataset.empty()) {
                // Twice the pair budget plus two: the retrieval corner (1 query x `budget` candidates) then fits unclamped.
                std::size_t const budget = std::stoull(env_or("STRINGWARS_BATCH_PER_CORE", "256")) * cores;
                return synthesize(alphabet, synthetic_length, std::max<std::size_t>(4096, 2 * budget + 2), settings.seed);
            }
            std::ifstream file(dataset, std::ios::binary);
            if (!file) throw std::runtime_error("cannot read " + dataset);
            std::ostringstream content;
            content << file.rdbuf();
            return tokenize(content.str(), tokens, settings.seed, dataset + " split by " + tokens);
        };

        {
            Corpus const corpus = load("ACGT");
            report("Levenshtein", corpus);
            bench_levenshtein(settings, corpus, make_shapes(corpus, cores), accelerated, baseline);
        }
        {
            Corpus const corpus = load("ARNDCQEGHILKMFPSTWYV");
            report("Needleman-Wunsch / Smith-Waterman", corpus);
            bench_needleman_wunsch_smith_waterman(settings, corpus, make_shapes(corpus, cores), accelerated, baseline);
        }
        std::printf("\nAll benchmarks finished.\n");
        std::fflush(stdout);
        // Leave without static destructors: a CUDA runtime tearing down beside a still-spinning CPU pool can take minutes.
        std::_Exit(0);
    }
    catch (std::exception const &error) {
        std::fprintf(stderr, "Failed with: %s\n", error.what());
        return 1;
    }
}
