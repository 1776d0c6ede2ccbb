"""CPU suite, part 4: the host scheduler (engine.cu `build_schedule`) on string LENGTHS alone, through
`szs_b200_debug_schedule` — no GPU involved. Invariants: every candidate lands in exactly one tier; shapes cover the
lengths; the jobs of a slot tile its query range exactly once; the query order is longest-first."""

import ctypes as ct

import numpy as np
import pytest

import stringzilla_b200 as szb
from stringzilla_b200 import _abi, scoring

TIER_BIT_PARALLEL, TIER_PACKED, TIER_WIDE = 0, 1, 2


def plan(engine, query_lengths, candidate_lengths, distinct_bytes=4, device_sms=148):
    lib = _abi.bind()
    q = np.ascontiguousarray(query_lengths, dtype=np.uint32)
    symmetric = candidate_lengths is None
    c = q if symmetric else np.ascontiguousarray(candidate_lengths, dtype=np.uint32)
    assignment = np.zeros(len(c), dtype=np.uint32)
    partner = np.zeros(len(c), dtype=np.uint32)
    order = np.zeros(len(q), dtype=np.uint32)
    capacity = 200000
    jobs = np.zeros(3 * capacity, dtype=np.uint32)
    n_jobs = ct.c_uint64()
    status = lib.szs_b200_debug_schedule(
        engine._handle, q.ctypes.data, len(q), None if symmetric else c.ctypes.data, len(c), distinct_bytes, device_sms,
        assignment.ctypes.data, partner.ctypes.data, order.ctypes.data, jobs.ctypes.data, capacity, ct.byref(n_jobs))
    assert status == 0
    assert n_jobs.value <= capacity
    return q, c, assignment, partner, order, jobs[: 3 * n_jobs.value].reshape(-1, 3)


def check_invariants(q, c, assignment, partner, order, jobs, symmetric=False):
    tier, width, lanes = assignment & 0xFF, (assignment >> 8) & 0xFF, (assignment >> 16) & 0xFF
    assert (assignment != 0xFFFFFFFF).all(), "every candidate belongs to exactly one tier"
    assert sorted(order.tolist()) == list(range(len(q)))
    assert (np.diff(q[order].astype(np.int64)) <= 0).all(), "query order is longest first"
    for cand in range(len(c)):
        assert lanes[cand] in (1, 2, 4, 8, 16, 32)
        if tier[cand] == TIER_BIT_PARALLEL:
            assert 32 * int(width[cand]) * int(lanes[cand]) >= c[cand], "pattern words cover the candidate"
        else:
            assert width[cand] % 4 == 0 and lanes[cand] >= 8
        if partner[cand] != 0xFFFFFFFF:
            assert partner[partner[cand]] == cand and assignment[partner[cand]] == assignment[cand]
            assert tier[cand] == TIER_PACKED, "only the packed tier shares registers between two candidates"
    # Jobs of one slot tile [0, rows) of the query order without gaps or overlaps.
    position = np.empty(len(q), dtype=np.int64)
    position[order] = np.arange(len(q))
    by_slot = {}
    for c0, begin, end in jobs.tolist():
        by_slot.setdefault(c0, []).append((begin, end))
    dp_first = {int(cand) for cand in range(len(c)) if tier[cand] != TIER_BIT_PARALLEL and
                (partner[cand] == 0xFFFFFFFF or cand in by_slot)}
    for c0, ranges in by_slot.items():
        ranges.sort()
        assert ranges[0][0] == 0 and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        rows = ranges[-1][1]
        members = [c0] + ([int(partner[c0])] if partner[c0] != 0xFFFFFFFF else [])
        needed = max(int(position[m]) + 1 for m in members) if symmetric else len(q)
        assert rows == needed, (c0, rows, needed)
    scheduled = set(by_slot) | {int(partner[c0]) for c0 in by_slot if partner[c0] != 0xFFFFFFFF}
    assert scheduled == {int(cand) for cand in range(len(c)) if tier[cand] != TIER_BIT_PARALLEL}
    return tier


def test_unit_cost_dna_takes_the_bit_parallel_tier_up_to_its_widest_shape():
    rng = np.random.default_rng(1)
    q = rng.integers(0, 3000, size=300)
    c = np.concatenate([rng.integers(0, 10240, size=200), [10240, 10241, 20000, 0, 1]])
    tier = check_invariants(*plan(szb.LevenshteinDistances(), q, c))
    assert ((tier == TIER_BIT_PARALLEL) == (c <= 10240)).all()
    assert (check_invariants(*plan(szb.LevenshteinDistances(bit_parallel=False), q, c)) != TIER_BIT_PARALLEL).all()


def test_cost_magnitudes_pick_the_cell_width():
    rng = np.random.default_rng(2)
    q, c = rng.integers(1, 900, size=64), rng.integers(1, 900, size=70)
    assert (check_invariants(*plan(szb.LevenshteinDistances(1, 3, 3, 3), q, c)) == TIER_PACKED).all()
    tier = check_invariants(*plan(szb.LevenshteinDistances(0, 100, 100, 100), q, c))
    assert ((tier == TIER_WIDE) == ((np.minimum(c, q.max()) + 2) * 200 > 65000)).all()
    byte_to_class, table = scoring.blosum62()
    for engine_type in (szb.NeedlemanWunschScores, szb.SmithWatermanScores):
        for gaps in ((-4, -4), (-4, -1)):
            tier = check_invariants(*plan(engine_type(byte_to_class, table, open=gaps[0], extend=gaps[1]), q, c, 24))
            assert (tier == TIER_PACKED).all()
        # Gap rewards break the packed recurrences' preconditions: literal 32-bit tier.
        assert (check_invariants(*plan(engine_type(byte_to_class, table, open=-2, extend=1), q, c, 24)) == TIER_WIDE).all()


@pytest.mark.parametrize("count", [1, 2, 33, 500])
def test_symmetric_calls_schedule_a_prefix_per_candidate(count):
    rng = np.random.default_rng(count)
    q = rng.integers(0, 1500, size=count)
    byte_to_class, table = scoring.blosum62()
    for engine in (szb.LevenshteinDistances(1, 3, 3, 3), szb.NeedlemanWunschScores(byte_to_class, table, open=-4, extend=-1)):
        check_invariants(*plan(engine, q, None, 24), symmetric=count > 1)


def test_extreme_shapes():
    check_invariants(*plan(szb.LevenshteinDistances(1, 3, 3, 3), [5], np.full(100000, 3)))
    check_invariants(*plan(szb.LevenshteinDistances(1, 3, 3, 3), np.full(50000, 7), [9]))
    check_invariants(*plan(szb.LevenshteinDistances(2, 3, 4, 1), [70000, 3], [65000, 123456, 0]))


def test_calls_too_small_to_fill_the_device_take_latency_shapes():
    """A call with fewer pairs than the device has warps is bound by the latency of one sweep: each pair is spread over
    as many lanes as there are idle ones (one word per lane for a single 1024-character pair, the widest DP wavefront)."""
    unit = szb.LevenshteinDistances()
    q, c, assignment, partner, order, jobs = plan(unit, [1024], [1024])
    tier, width, lanes = assignment & 0xFF, (assignment >> 8) & 0xFF, (assignment >> 16) & 0xFF
    assert tier[0] == TIER_BIT_PARALLEL and lanes[0] == 32 and width[0] == 1
    # c1: 32 x 32 pairs of 64..256 characters still leave most warps idle: at least 8 lanes per pair.
    rng = np.random.default_rng(5)
    q, c, assignment, partner, order, jobs = plan(unit, rng.integers(64, 257, size=32), rng.integers(64, 257, size=32))
    assert ((assignment & 0xFF) == TIER_BIT_PARALLEL).all() and (((assignment >> 16) & 0xFF) >= 8).all()
    check_invariants(q, c, assignment, partner, order, jobs)
    # The c2 tile fills the device: the tightest coverage wins again (4 lanes of 8 or 9 words).
    q, c, assignment, partner, order, jobs = plan(unit, rng.integers(960, 1089, size=316), rng.integers(960, 1089, size=317))
    assert set(((assignment >> 16) & 0xFF).tolist()) == {4} and set(((assignment >> 8) & 0xFF).tolist()) <= {8, 9}
    # DP tier of a tiny call: the widest wavefront (32 lanes).
    weird = szb.LevenshteinDistances(1, 3, 3, 3)
    q, c, assignment, partner, order, jobs = plan(weird, [700, 650], [600, 640, 100])
    assert ((assignment & 0xFF) == TIER_PACKED).all() and (((assignment >> 16) & 0xFF) == 32).all()
    check_invariants(q, c, assignment, partner, order, jobs)


def test_jobs_are_single_rounds_on_large_tiles():
    """Jobs are cut to ~24 per SM so that the ragged end of a persistent launch costs a few percent: on the c3 tile every
    job is one round of the CTA's sixteen warps."""
    byte_to_class, table = scoring.blosum62()
    engine = szb.NeedlemanWunschScores(byte_to_class, table, open=-4, extend=-1)
    rng = np.random.default_rng(6)
    q, c, assignment, partner, order, jobs = plan(engine, rng.integers(960, 1089, size=316), rng.integers(960, 1089, size=317),
                                                  distinct_bytes=20)
    check_invariants(q, c, assignment, partner, order, jobs)
    assert len(jobs) >= 148 * 16 and max(end - begin for _, begin, end in jobs.tolist()) == 16


def _engine_zoo():
    byte_to_class, table = scoring.blosum62()
    return [
        szb.LevenshteinDistances(),
        szb.LevenshteinDistances(bit_parallel=False),
        szb.LevenshteinDistances(1, 3, 3, 3),
        szb.LevenshteinDistances(1, 3, 4, 2),
        szb.LevenshteinDistances(0, 100, 100, 100),
        szb.NeedlemanWunschScores(byte_to_class, table, open=-4, extend=-4),
        szb.NeedlemanWunschScores(byte_to_class, table, open=-4, extend=-1),
        szb.SmithWatermanScores(byte_to_class, table, open=-4, extend=-1),
        szb.SmithWatermanScores(byte_to_class, table, open=-2, extend=1),
    ]


def _random_lengths(rng, count, longest, flavour):
    if flavour == 0:
        return rng.integers(0, longest + 1, size=count)
    if flavour == 1:  # log-uniform, like the c4 tile
        return np.exp(rng.uniform(0, np.log(longest + 1), size=count)).astype(np.int64)
    if flavour == 2:  # a narrow band around one length, like c2
        return rng.integers(max(0, longest - longest // 8), longest + 1, size=count)
    lengths = rng.integers(0, longest + 1, size=count)  # many empty and equal strings
    lengths[rng.random(count) < 0.4] = 0
    lengths[rng.random(count) < 0.3] = longest
    return lengths


@pytest.mark.parametrize("seed", range(60))
def test_random_calls_keep_the_schedule_invariants(seed):
    """Randomised sweep over engines, length distributions, alphabets, device sizes and symmetric calls: the invariants of
    `check_invariants` hold for every schedule the host builds (the scheduler is pure host code, no GPU involved)."""
    rng = np.random.default_rng(1000 + seed)
    engines = _engine_zoo()
    engine = engines[seed % len(engines)]
    n_queries = int(rng.choice([1, 2, 7, 33, 200, 1500]))
    n_candidates = int(rng.choice([1, 3, 40, 250]))
    longest = int(rng.choice([0, 1, 31, 300, 1100, 5000, 12000]))
    flavour = int(rng.integers(0, 4))
    distinct = int(rng.choice([1, 4, 24, 31, 32, 200]))
    device_sms = int(rng.choice([1, 8, 148, 160]))
    symmetric = bool(rng.integers(0, 3) == 0)
    q = _random_lengths(rng, n_queries, longest, flavour)
    c = None if symmetric else _random_lengths(rng, n_candidates, longest, int(rng.integers(0, 4)))
    result = plan(engine, q, c, distinct_bytes=distinct, device_sms=device_sms)
    check_invariants(*result, symmetric=symmetric and n_queries > 1)


def test_structural_efficiency_of_the_bench_tiles():
    """Useful cells / executed cells (tools/structural_efficiency.py: padding up to a shape's coverage, wavefront fill and
    drain, the shorter partner of a packed pair, multi-pass rounding) for the BASELINE tiles on a 148-SM device: a guard
    against a scheduler change that quietly wastes cells. The measured values are in DESIGN.md §7."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("structural_efficiency", support_root() / "tools" / "structural_efficiency.py")
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    for workload, floor in (("c2", 0.93), ("c2_weird", 0.90), ("c3_sw", 0.90), ("c4", 0.88)):
        label, value = module.efficiency(workload)
        assert floor <= value <= 1.0, (label, value)


def support_root():
    from tests import support

    return support.ROOT
