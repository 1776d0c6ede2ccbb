"""GPU suite: the CUDA path, called through the C ABI, must equal the oracle bit for bit (all three families are exact
integer DP — the reference's own tests assert equality, test/similarities.cuh:543-545, 676-680, 1229-1232)."""

import ctypes as ct
import json

import numpy as np
import pytest

import stringzilla_b200 as szb
from stringzilla_b200 import _abi, scoring
from tests import support

pytestmark = pytest.mark.gpu

GOLDEN = json.loads((support.GOLDEN_DIR / "reference_vectors.json").read_text())


@pytest.fixture(scope="module")
def scope():
    return szb.DeviceScope(gpu_device=0)


def engine_for(kind, params):
    if kind == "levenshtein":
        return szb.LevenshteinDistances(*params["costs"])
    if kind == "levenshtein_utf8":
        return szb.LevenshteinDistancesUTF8(*params["costs"])
    engine_type = szb.NeedlemanWunschScores if kind == "needleman_wunsch" else szb.SmithWatermanScores
    return engine_type(np.array(params["byte_to_class"], dtype=np.uint8),
                       np.array(params["costs"], dtype=np.int8).reshape(32, 32),
                       open=params["open"], extend=params["extend"])  # fmt: skip


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=[c["name"] for c in GOLDEN["cases"]])
def test_golden_vectors_of_the_reference(case, scope):
    queries = [bytes.fromhex(q) for q in case["queries"]]
    candidates = None if case["candidates"] is None else [bytes.fromhex(c) for c in case["candidates"]]
    got = engine_for(case["kind"], case["params"])(queries, candidates, device=scope)
    assert np.array_equal(got.astype(np.int64), np.array(case["matrix"], dtype=np.int64))


def test_known_answer_levenshtein_vectors(scope):
    engine = szb.LevenshteinDistances()
    for first, second, distance in GOLDEN["known_levenshtein"]:
        assert int(engine([first.encode()], [second.encode()], device=scope)[0, 0]) == distance, (first, second)
    firsts = [a.encode() for a, _, _ in GOLDEN["known_levenshtein"]]
    seconds = [b.encode() for _, b, _ in GOLDEN["known_levenshtein"]]
    batch = engine(firsts, seconds, device=scope)
    assert [int(batch[i, i]) for i in range(len(firsts))] == [d for _, _, d in GOLDEN["known_levenshtein"]]


# Length bands chosen to land in every strip shape: (8,8) <=64, (16,8) <=128, (32,8) <=256, (32,16) <=512, (32,32)
# single pass <=1024 and multi-pass beyond.
LENGTH_BANDS = [(1, 64), (50, 128), (100, 256), (200, 512), (400, 1024), (900, 2600)]
LEVENSHTEIN_COSTS = [(0, 1, 1, 1), (1, 3, 3, 3), (1, 3, 4, 2), (0, 2, 5, 1)]


@pytest.mark.parametrize("band", LENGTH_BANDS, ids=[f"len{a}-{b}" for a, b in LENGTH_BANDS])
@pytest.mark.parametrize("costs", LEVENSHTEIN_COSTS, ids=["unit", "weird_linear", "weird_affine", "affine"])
def test_levenshtein_random_cross_products(band, costs, scope):
    rng = np.random.default_rng(hash((band, costs)) & 0xFFFF)
    n = 23 if band[1] <= 1024 else 7
    queries = support.random_strings(rng, n, band[0], band[1], support.DNA)
    candidates = support.mutated_copies(rng, queries[: n - 4], 0.15, support.DNA) + support.random_strings(
        rng, 5, band[0], band[1], support.DNA)
    want = support.oracle_levenshtein(queries, candidates, *costs)
    assert np.array_equal(szb.LevenshteinDistances(*costs)(queries, candidates, device=scope), want)
    # Unit costs normally take the bit-parallel kernel; the DPX wavefront tier must agree with it and the oracle.
    assert np.array_equal(szb.LevenshteinDistances(*costs, bit_parallel=False)(queries, candidates, device=scope), want)


def test_bit_parallel_levenshtein_shapes(scope):
    """Every (words per lane, lanes) shape of the bit-parallel kernel: pattern lengths 0..9000 (longer ones fall back to
    the DP tiers), ragged text lengths, a 21-letter alphabet, and the engine statistics proving which kernel ran."""
    rng = np.random.default_rng(29)
    lengths = [0, 1, 31, 32, 33, 63, 65, 100, 129, 255, 256, 257, 511, 513, 767, 769, 1023, 1024, 1025, 1279, 1281,
               1535, 1537, 1791, 1793, 2047, 2049, 3071, 3073, 4095, 4097, 5000, 6143, 6145, 7167, 7169, 8191, 8192,
               8193, 9000, 12289, 16383, 16384, 16385, 17000]
    candidates = [support.random_strings(rng, 1, n, n, support.DNA)[0] for n in lengths]
    # 33 queries: fewer than 32 rows against more candidates would trade roles (and turn everything bit-parallel)
    queries = support.random_strings(rng, 30, 0, 3000, support.DNA) + [b"", candidates[17][::-1], candidates[9]]
    engine = szb.LevenshteinDistances()
    got = engine(queries, candidates, device=scope)
    assert np.array_equal(got, support.oracle_levenshtein(queries, candidates))
    stats = engine.last_call_stats
    assert stats.pairs_bit_parallel > 0 and stats.pairs_u16 > 0  # 16385 and 17000 exceed the widest bit-parallel shape
    proteins = support.random_strings(rng, 9, 0, 700, support.PROTEIN)
    others = support.mutated_copies(rng, proteins, 0.2, support.PROTEIN) + support.random_strings(rng, 7, 0, 1500, support.PROTEIN)
    engine = szb.LevenshteinDistances()
    assert np.array_equal(engine(proteins, others, device=scope), support.oracle_levenshtein(proteins, others))
    assert engine.last_call_stats.pairs_bit_parallel > 0
    # Symmetric mode and many short strings (several queries per warp).
    shorts = support.random_strings(rng, 70, 0, 200, support.DNA)
    assert np.array_equal(szb.LevenshteinDistances()(shorts, device=scope), support.oracle_levenshtein(shorts, None))


@pytest.mark.parametrize("band", LENGTH_BANDS, ids=[f"len{a}-{b}" for a, b in LENGTH_BANDS])
@pytest.mark.parametrize("gaps", [(-4, -4), (-4, -1), (-11, -1)], ids=["linear", "affine", "affine_wide"])
@pytest.mark.parametrize("local", [False, True], ids=["needleman_wunsch", "smith_waterman"])
def test_weighted_random_cross_products(band, gaps, local, scope):
    rng = np.random.default_rng(hash((band, gaps, local)) & 0xFFFF)
    byte_to_class, costs = scoring.blosum62()
    n = 19 if band[1] <= 1024 else 6
    queries = support.random_strings(rng, n, band[0], band[1], support.PROTEIN)
    candidates = support.mutated_copies(rng, queries[: n - 3], 0.25, support.PROTEIN) + support.random_strings(
        rng, 4, band[0], band[1], support.PROTEIN)
    engine_type = szb.SmithWatermanScores if local else szb.NeedlemanWunschScores
    got = engine_type(byte_to_class, costs, open=gaps[0], extend=gaps[1])(queries, candidates, device=scope)
    assert np.array_equal(got, support.oracle_weighted(queries, candidates, byte_to_class, costs, *gaps, local))


def test_mixed_lengths_one_call(scope):
    """Log-uniform lengths 1..3000 in one tile: every strip shape, both tiers and multi-pass slots in one launch set."""
    rng = np.random.default_rng(11)
    lengths = np.exp(rng.uniform(0, np.log(3000), size=40)).astype(int)
    strings = [support.random_strings(rng, 1, int(n), int(n), support.PROTEIN)[0] for n in lengths]
    queries, candidates = strings[:17], strings[17:] + [b""]
    byte_to_class, costs = scoring.blosum62()
    for engine_type, local in ((szb.NeedlemanWunschScores, False), (szb.SmithWatermanScores, True)):
        got = engine_type(byte_to_class, costs, open=-4, extend=-1)(queries, candidates, device=scope)
        assert np.array_equal(got, support.oracle_weighted(queries, candidates, byte_to_class, costs, -4, -1, local))
    got = szb.LevenshteinDistances()(queries, candidates, device=scope)
    assert np.array_equal(got, support.oracle_levenshtein(queries, candidates))


def test_shapes_and_degenerate_inputs(scope):
    """1xN, Nx1, 1x1, ragged with empties, Nx0 / 0xN / 0x0 (reference test/similarities.cuh:1187-1326)."""
    rng = np.random.default_rng(3)
    engine = szb.LevenshteinDistances()
    pool = support.random_strings(rng, 9, 0, 40, b"ab") + [b"", b""]
    for queries, candidates in ((pool[:1], pool), (pool, pool[:1]), (pool[:1], pool[1:2]), (pool, pool[::-1]),
                                ([b""], [b""]), ([b""], [b"xyz"]), ([b"xyz"], [b""])):
        got = engine(queries, candidates, device=scope)
        assert np.array_equal(got, support.oracle_levenshtein(queries, candidates))
    assert engine([], pool, device=scope).shape == (0, len(pool))
    assert engine(pool, [], device=scope).shape == (len(pool), 0)
    assert engine([], [], device=scope).shape == (0, 0)
    sample = b"ADVERSARIAL"
    for gap in (1, 3):
        tuned = szb.LevenshteinDistances(0, 1, gap, gap)
        assert int(tuned([sample], [b""], device=scope)[0, 0]) == len(sample) * gap
        assert int(tuned([sample], [sample], device=scope)[0, 0]) == 0


def test_symmetric_self_similarity(scope):
    rng = np.random.default_rng(5)
    strings = support.random_strings(rng, 21, 0, 300, support.DNA)
    matrix = szb.LevenshteinDistances()(strings, device=scope)
    assert np.array_equal(matrix, support.oracle_levenshtein(strings, None))
    assert np.array_equal(matrix, matrix.T) and not matrix.diagonal().any()
    byte_to_class, costs = scoring.blosum62()
    proteins = support.random_strings(rng, 15, 1, 200, support.PROTEIN)
    for engine_type, local in ((szb.NeedlemanWunschScores, False), (szb.SmithWatermanScores, True)):
        got = engine_type(byte_to_class, costs, open=-4, extend=-1)(proteins, device=scope)
        assert np.array_equal(got, support.oracle_weighted(proteins, None, byte_to_class, costs, -4, -1, local))


def test_symmetric_triangle_scheduling(scope):
    """Symmetric calls under symmetric costs score each unordered pair once (longer string as the query) and mirror by
    rank; equal lengths, empties, multi-pass lengths and both Levenshtein kernels must all survive that."""
    rng = np.random.default_rng(31)
    strings = (support.random_strings(rng, 30, 0, 400, support.DNA) + [b"", b"", b"ACGT" * 50, b"ACGT" * 50]
               + support.random_strings(rng, 6, 1400, 3000, support.DNA))
    for costs, kwargs in (((0, 1, 1, 1), {}), ((0, 1, 1, 1), {"bit_parallel": False}), ((1, 3, 4, 2), {})):
        got = szb.LevenshteinDistances(*costs, **kwargs)(strings, device=scope)
        assert np.array_equal(got, support.oracle_levenshtein(strings, None, *costs))
    byte_to_class, costs = scoring.blosum62()
    proteins = support.random_strings(rng, 25, 0, 500, support.PROTEIN) + support.random_strings(rng, 3, 1300, 2500, support.PROTEIN)
    for engine_type, local in ((szb.NeedlemanWunschScores, False), (szb.SmithWatermanScores, True)):
        for gaps in ((-4, -4), (-4, -1)):
            got = engine_type(byte_to_class, costs, open=gaps[0], extend=gaps[1])(proteins, device=scope)
            assert np.array_equal(got, support.oracle_weighted(proteins, None, byte_to_class, costs, *gaps, local))


def test_asymmetric_substitution_table(scope):
    """cost[class(query)][class(candidate)] — the query indexes the row (SURVEY §7 'hard parts')."""
    rng = np.random.default_rng(8)
    byte_to_class, costs = scoring.uniform(b"ACGT", 2, -1)
    costs = costs.copy()
    costs[1, 2], costs[2, 1], costs[3, 4], costs[4, 3] = -3, 1, 0, -2
    queries = support.random_strings(rng, 11, 1, 300, support.DNA)
    candidates = support.mutated_copies(rng, queries, 0.2, support.DNA)
    for engine_type, local in ((szb.NeedlemanWunschScores, False), (szb.SmithWatermanScores, True)):
        for gaps in ((-2, -2), (-3, -1)):
            got = engine_type(byte_to_class, costs, open=gaps[0], extend=gaps[1])(queries, candidates, device=scope)
            assert np.array_equal(got, support.oracle_weighted(queries, candidates, byte_to_class, costs, *gaps, local))
        got = engine_type(byte_to_class, costs, open=-2, extend=-1)(queries, device=scope)
        assert np.array_equal(got, support.oracle_weighted(queries, None, byte_to_class, costs, -2, -1, local))


def test_bit_parallel_levenshtein_large_alphabets(scope):
    """More than 31 distinct bytes: the CTA-shared match table ranks the bytes each candidate contains."""
    rng = np.random.default_rng(41)
    text = bytes(range(32, 127))
    every_byte = bytes(range(256))
    for alphabet, top in ((text, 1400), (every_byte, 900), (every_byte, 5000)):
        lengths = [0, 1, 33, 200, 513, 1024, 1025, top]
        candidates = [support.random_strings(rng, 1, n, n, alphabet)[0] for n in lengths]
        queries = support.random_strings(rng, 21, 0, top, alphabet) + [b"", candidates[4], candidates[5][:-3] + b"xyz"]
        engine = szb.LevenshteinDistances()
        got = engine(queries, candidates, device=scope)
        assert np.array_equal(got, support.oracle_levenshtein(queries, candidates))
        assert engine.last_call_stats.pairs_bit_parallel > 0
    strings = support.random_strings(rng, 40, 0, 300, text)
    assert np.array_equal(szb.LevenshteinDistances()(strings, device=scope), support.oracle_levenshtein(strings, None))


def test_large_alphabets_and_wide_cells(scope):
    """More than 31 distinct bytes -> byte-compare recurrences; high costs -> 32-bit cells (test/similarities.cuh:2023)."""
    rng = np.random.default_rng(13)
    every_byte = bytes(range(256))
    queries = support.random_strings(rng, 9, 0, 700, every_byte)
    candidates = support.mutated_copies(rng, queries, 0.3, every_byte)
    for costs in ((0, 1, 1, 1), (1, 3, 4, 2), (1, 3, 3, 3), (0, 5, 2, 2), (2, 1, 3, 3)):
        engine = szb.LevenshteinDistances(*costs, bit_parallel=False)
        got = engine(queries, candidates, device=scope)
        assert np.array_equal(got, support.oracle_levenshtein(queries, candidates, *costs)), costs
        # Linear gaps with match <= mismatch take the packed byte-compare recurrence; the rest the 32-bit tier.
        packed = costs[2] == costs[3] and costs[0] <= costs[1]
        assert (engine.last_call_stats.pairs_u16 > 0) == packed, costs
    binary_q = support.random_strings(rng, 5, 600, 800, b"ab")
    binary_c = support.random_strings(rng, 6, 600, 800, b"ab")
    for costs in ((0, 100, 100, 100), (0, 127, 90, 60)):
        engine = szb.LevenshteinDistances(*costs)
        got = engine(binary_q, binary_c, device=scope)
        assert np.array_equal(got, support.oracle_levenshtein(binary_q, binary_c, *costs))
        assert engine.last_call_stats.pairs_s32 > 0
    byte_to_class, table = scoring.uniform(b"ab", 100, -100)
    got = szb.NeedlemanWunschScores(byte_to_class, table, open=-100, extend=-100)(binary_q, binary_c, device=scope)
    assert np.array_equal(got, support.oracle_weighted(binary_q, binary_c, byte_to_class, table, -100, -100, False))


def test_unusual_cost_signs_take_the_literal_tier(scope):
    """Gap rewards and opening cheaper than extending defeat the packed recurrences' preconditions."""
    rng = np.random.default_rng(17)
    byte_to_class, table = scoring.blosum62()
    queries = support.random_strings(rng, 7, 0, 150, support.PROTEIN)
    candidates = support.random_strings(rng, 6, 0, 150, support.PROTEIN)
    for engine_type, local in ((szb.NeedlemanWunschScores, False), (szb.SmithWatermanScores, True)):
        for gaps in ((-1, -4), (-2, 1), (1, 1), (3, -2)):
            got = engine_type(byte_to_class, table, open=gaps[0], extend=gaps[1])(queries, candidates, device=scope)
            want = support.oracle_weighted(queries, candidates, byte_to_class, table, *gaps, local)
            assert np.array_equal(got, want), (engine_type.__name__, gaps)


def test_utf8_levenshtein_rune_level(scope):
    """Rune-level distances (reference test/similarities.py:231-260): accents, CJK, emoji; random multi-script text at
    lengths that need several strip shapes; pure ASCII must equal the byte-level engine."""
    engine = szb.LevenshteinDistancesUTF8()
    distance = lambda a, b: int(engine([a.encode()], [b.encode()], device=scope)[0, 0])
    assert distance("hello", "hell") == 1
    assert distance("𠜎 𠜱 𠝹 𠱓", "𠜎𠜱𠝹𠱓") == 3
    assert distance("💖", "💗") == 1
    assert distance("αβγδ", "αγδ") == 1
    assert distance("façade", "facade") == 1
    assert distance("Schön", "Scho\u0308n") == 2
    assert distance("こんにちは世界", "こんばんは世界") == 2
    rng = np.random.default_rng(37)
    alphabet = list("abcdefgh абвгдежз αβγδεζηθ 数据科学机器 🙂🌍🚀💖")
    def text(n):
        return "".join(alphabet[i] for i in rng.integers(0, len(alphabet), size=n)).encode()
    queries = [text(int(n)) for n in rng.integers(0, 900, size=12)] + [b"", "é".encode()]
    candidates = [text(int(n)) for n in rng.integers(0, 1500, size=9)] + [b"", b"plain"]
    for costs in ((0, 1, 1, 1), (1, 3, 3, 3), (0, 4, 3, 2)):
        got = szb.LevenshteinDistancesUTF8(*costs)(queries, candidates, device=scope)
        assert np.array_equal(got, support.oracle_levenshtein_utf8(queries, candidates, *costs)), costs
    got = szb.LevenshteinDistancesUTF8()(queries, device=scope)
    assert np.array_equal(got, support.oracle_levenshtein_utf8(queries, None))
    # More than 256 distinct runes in one call: no byte ranking exists. Unit costs take the rune-keyed bit-parallel kernel
    # (candidates of up to 1024 runes; the 1500-rune one stays on the 32-bit symbol recurrences), other costs the
    # 32-bit symbol recurrences.
    cjk = [chr(0x4E00 + i) for i in range(2000)]
    wide_text = lambda n: "".join(cjk[i] for i in rng.integers(0, len(cjk), size=n)).encode()
    wide_q = [wide_text(int(n)) for n in rng.integers(0, 500, size=7)] + [text(40), wide_text(1024), wide_text(1100)]
    wide_c = [wide_text(int(n)) for n in rng.integers(0, 700, size=6)] + [wide_q[0][: 3 * 50], b"ascii too", wide_text(1),
                                                                          wide_text(33), wide_text(1024), wide_text(1500)]
    for costs in ((0, 1, 1, 1), (0, 4, 3, 2)):
        engine = szb.LevenshteinDistancesUTF8(*costs)
        assert np.array_equal(engine(wide_q, wide_c, device=scope), support.oracle_levenshtein_utf8(wide_q, wide_c, *costs))
        stats = engine.last_call_stats
        assert stats.pairs_s32 > 0 and (stats.pairs_bit_parallel > 0) == (costs == (0, 1, 1, 1))
    # The rune-keyed kernel over every wavefront width (1 .. 32 lanes of one 32-bit word), symmetric call included.
    many_c = [wide_text(n) for n in (1, 31, 32, 33, 64, 65, 128, 129, 256, 257, 512, 513, 1000, 1024)]
    many_q = [wide_text(int(n)) for n in rng.integers(0, 1200, size=40)]
    rune_engine = szb.LevenshteinDistancesUTF8()
    assert np.array_equal(rune_engine(many_q, many_c, device=scope), support.oracle_levenshtein_utf8(many_q, many_c))
    assert rune_engine.last_call_stats.pairs_s32 == 0
    assert np.array_equal(rune_engine(many_c, device=scope), support.oracle_levenshtein_utf8(many_c, None))
    few = szb.LevenshteinDistancesUTF8()
    assert np.array_equal(few(queries, candidates, device=scope), support.oracle_levenshtein_utf8(queries, candidates))
    assert few.last_call_stats.pairs_bit_parallel > 0  # ~40 distinct runes: ranked into bytes, bit-parallel kernel
    ascii_q = support.random_strings(rng, 8, 0, 300, b"abcdefghij ")
    ascii_c = support.random_strings(rng, 7, 0, 300, b"abcdefghij ")
    rune_level = szb.LevenshteinDistancesUTF8()
    assert np.array_equal(rune_level(ascii_q, ascii_c, device=scope), support.oracle_levenshtein(ascii_q, ascii_c))
    assert rune_level.last_call_stats.pairs_bit_parallel > 0  # pure ASCII takes the byte-level fast path


def test_every_input_layout_of_the_abi(scope):
    """u32 tapes, u64 tapes and callback sequences; a row stride wider than the matrix; `out=` reuse."""
    rng = np.random.default_rng(19)
    queries = support.random_strings(rng, 6, 0, 90, support.DNA)
    candidates = support.random_strings(rng, 5, 0, 90, support.DNA)
    want = support.oracle_levenshtein(queries, candidates)
    lib = _abi.bind()
    engine = szb.LevenshteinDistances()
    q, c = szb.Tape(queries), szb.Tape(candidates)

    wide = np.full((len(queries), 9), 0xDEAD, dtype=np.uint64)
    out = engine(q, c, device=scope, out=wide[:, :5])
    assert np.array_equal(out, want) and (wide[:, 5:] == 0xDEAD).all()

    q32, c32 = q.offsets.astype(np.uint32), c.offsets.astype(np.uint32)
    tq = _abi.SequenceU32Tape(q.data.ctypes.data, q32.ctypes.data, len(q))
    tc = _abi.SequenceU32Tape(c.data.ctypes.data, c32.ctypes.data, len(c))
    got = np.zeros_like(want)
    message = ct.c_char_p()
    status = lib.szs_levenshtein_distances_u32tape(engine._handle, scope.handle, ct.byref(tq), ct.byref(tc),
                                                   got.ctypes.data, 5, ct.byref(message))
    assert status == 0 and np.array_equal(got, want)

    def sequence_of(strings):
        buffers = [ct.create_string_buffer(s, len(s)) if s else ct.create_string_buffer(1) for s in strings]
        start = _abi.SEQUENCE_START(lambda _, i: ct.addressof(buffers[i]))
        length = _abi.SEQUENCE_LENGTH(lambda _, i: len(strings[i]))
        return _abi.Sequence(None, len(strings), start, length), (buffers, start, length)

    sq, keep_q = sequence_of(queries)
    sc, keep_c = sequence_of(candidates)
    got = np.zeros_like(want)
    status = lib.szs_levenshtein_distances(engine._handle, scope.handle, ct.byref(sq), ct.byref(sc), got.ctypes.data,
                                           5, ct.byref(message))
    assert status == 0 and np.array_equal(got, want)


@pytest.mark.parametrize("memory", ["unified_memory", "cuda_memory"])
def test_callback_sequences_outside_host_memory(memory, scope):
    """Callback sequences whose strings live in managed memory (the reference bench's `unified_alloc` tokens) and in device
    memory: the strings are gathered on the device from an uploaded pointer list, the host never reads them."""
    rng = np.random.default_rng(19)
    queries = support.random_strings(rng, 6, 0, 90, support.DNA)
    candidates = support.random_strings(rng, 5, 0, 90, support.DNA)
    want = support.oracle_levenshtein(queries, candidates)
    lib = _abi.bind()
    engine = szb.LevenshteinDistances()
    message = ct.c_char_p()

    def sequence_in(strings, base_address):
        offsets = np.concatenate([[0], np.cumsum([len(s) for s in strings])]).astype(np.int64)
        start = _abi.SEQUENCE_START(lambda _, i: base_address + int(offsets[i]))
        length = _abi.SEQUENCE_LENGTH(lambda _, i: len(strings[i]))
        return _abi.Sequence(None, len(strings), start, length), (start, length, offsets)

    blob = b"".join(queries) + b"".join(candidates)
    split = sum(len(s) for s in queries)
    if memory == "unified_memory":
        base = lib.szs_unified_alloc(len(blob))
        ct.memmove(base, blob, len(blob))
    else:
        device_blob = support.DeviceBuffer(np.frombuffer(blob, dtype=np.uint8))
        base = device_blob.pointer
    sq, keep_q = sequence_in(queries, base)
    sc, keep_c = sequence_in(candidates, base + split)
    got = np.zeros_like(want)
    status = lib.szs_levenshtein_distances(engine._handle, scope.handle, ct.byref(sq), ct.byref(sc), got.ctypes.data,
                                           5, ct.byref(message))
    assert status == 0 and np.array_equal(got, want)
    if memory == "unified_memory":
        lib.szs_unified_free(base, len(blob))


@pytest.mark.parametrize("seed", range(12))
def test_randomized_differential(seed, scope):
    """Seeded random configurations: family, costs (incl. unusual signs), alphabet size, length distribution, symmetric
    or rectangular shape. Each seed draws several small tiles and compares every cell with the oracle."""
    rng = np.random.default_rng(1000 + seed)
    alphabets = [support.DNA, support.PROTEIN, bytes(range(97, 123)), bytes(range(32, 127)), bytes(range(256)), b"ab"]
    for _ in range(6):
        alphabet = alphabets[int(rng.integers(0, len(alphabets)))]
        top = int(rng.choice([8, 40, 150, 400, 1300]))
        n_q, n_c = int(rng.integers(1, 14)), int(rng.integers(1, 14))
        queries = support.random_strings(rng, n_q, 0, top, alphabet)
        symmetric = rng.random() < 0.25
        if symmetric:
            candidates = None
        elif rng.random() < 0.5:
            candidates = support.mutated_copies(rng, (queries * 3)[:n_c], 0.2, alphabet)
        else:
            candidates = support.random_strings(rng, n_c, 0, top, alphabet)
        family = int(rng.integers(0, 3))
        if family == 0:
            match, mismatch = int(rng.integers(0, 3)), int(rng.integers(1, 9))
            open_ = int(rng.integers(1, 9))
            extend = open_ if rng.random() < 0.5 else int(rng.integers(1, 9))
            got = szb.LevenshteinDistances(match, mismatch, open_, extend)(queries, candidates, device=scope)
            want = support.oracle_levenshtein(queries, candidates, match, mismatch, open_, extend)
            label = ("levenshtein", match, mismatch, open_, extend)
        else:
            n_symbols = min(len(alphabet), 31)
            byte_to_class = np.zeros(256, dtype=np.uint8)
            byte_to_class[np.frombuffer(alphabet, dtype=np.uint8)] = (np.arange(len(alphabet)) % n_symbols) + 1
            costs = rng.integers(-9, 12, size=(32, 32)).astype(np.int8)
            if rng.random() < 0.6:
                costs = np.minimum(costs, costs.T)  # symmetric tables take the triangle path in symmetric calls
            open_ = int(rng.integers(-12, 1)) if rng.random() < 0.85 else int(rng.integers(-3, 4))
            extend = open_ if rng.random() < 0.4 else (int(rng.integers(-6, 1)) if rng.random() < 0.85 else int(rng.integers(-3, 4)))
            local = family == 2
            engine_type = szb.SmithWatermanScores if local else szb.NeedlemanWunschScores
            got = engine_type(byte_to_class, costs, open=open_, extend=extend)(queries, candidates, device=scope)
            want = support.oracle_weighted(queries, candidates, byte_to_class, costs, open_, extend, local)
            label = ("smith_waterman" if local else "needleman_wunsch", open_, extend)
        assert np.array_equal(got, want), (label, top, len(alphabet), symmetric)


def test_c_program_against_the_reference_header(scope):
    """oracle/_ref/dropin_probe was compiled (in the build container) against the reference's own stringzillas.h and
    linked against libstringzillas_b200.so; here it must compute d("kitten", "sitting") = 3 on the GPU."""
    import subprocess

    binary = support.ROOT / "oracle" / "_ref" / "dropin_probe"
    if not binary.exists():
        pytest.skip("drop-in probe was not built (no reference tree at build time)")
    ran = subprocess.run([str(binary)], capture_output=True, text=True)
    assert ran.returncode == 0 and "status 0 distance 3" in ran.stdout, ran.stdout + ran.stderr


def test_memory_kinds_and_tape_slices(scope):
    """Inputs and outputs in device memory, in managed memory (`szs_unified_alloc`, what the reference's Python binding
    uses), and tapes whose first offset is not zero (a slice of a larger Arrow array)."""
    rng = np.random.default_rng(53)
    strings = support.random_strings(rng, 14, 0, 200, support.PROTEIN)
    whole = szb.Tape(strings)
    byte_to_class, table = scoring.blosum62()
    engine = szb.NeedlemanWunschScores(byte_to_class, table, open=-4, extend=-1)
    lib = _abi.bind()
    message = ct.c_char_p()
    # Queries = strings[3:9], candidates = strings[5:14], both as windows into the SAME blob with non-zero first offsets.
    q_offsets, c_offsets = whole.offsets[3:10].copy(), whole.offsets[5:15].copy()
    want = support.oracle_weighted(strings[3:9], strings[5:14], byte_to_class, table, -4, -1, False)

    def call(data_pointer, out_pointer):
        tq = _abi.SequenceU64Tape(data_pointer, q_offsets.ctypes.data, 6)
        tc = _abi.SequenceU64Tape(data_pointer, c_offsets.ctypes.data, 9)
        status = lib.szs_needleman_wunsch_scores_u64tape(engine._handle, scope.handle, ct.byref(tq), ct.byref(tc),
                                                         out_pointer, 9, ct.byref(message))
        assert status == 0, message.value

    host_out = np.zeros((6, 9), dtype=np.int64)
    call(whole.data.ctypes.data, host_out.ctypes.data)
    assert np.array_equal(host_out, want)

    device_data = support.DeviceBuffer(whole.data)
    device_out = support.DeviceBuffer(np.zeros((6, 9), dtype=np.int64))
    call(device_data.pointer, device_out.pointer)
    assert np.array_equal(device_out.numpy(), want)
    # Offsets in device memory as well: the host reads them through a bounce copy.
    device_q_offsets, device_c_offsets = support.DeviceBuffer(q_offsets), support.DeviceBuffer(c_offsets)
    tq = _abi.SequenceU64Tape(device_data.pointer, device_q_offsets.pointer, 6)
    tc = _abi.SequenceU64Tape(device_data.pointer, device_c_offsets.pointer, 9)
    again = support.DeviceBuffer(np.zeros((6, 9), dtype=np.int64))
    assert lib.szs_needleman_wunsch_scores_u64tape(engine._handle, scope.handle, ct.byref(tq), ct.byref(tc), again.pointer, 9,
                                                   ct.byref(message)) == 0, message.value
    assert np.array_equal(again.numpy(), want)

    managed = lib.szs_unified_alloc(int(whole.data.size))
    managed_out = lib.szs_unified_alloc(6 * 9 * 8)
    ct.memmove(managed, whole.data.ctypes.data, int(whole.data.size))
    call(managed, managed_out)
    got = np.frombuffer((ct.c_int64 * 54).from_address(managed_out), dtype=np.int64).reshape(6, 9)
    assert np.array_equal(got, want)
    lib.szs_unified_free(managed, int(whole.data.size))
    lib.szs_unified_free(managed_out, 6 * 9 * 8)


def test_arrow_arrays_as_inputs(scope):
    """Arrow string arrays (and slices of them) go in as they are; same matrix as Python lists (reference
    test/similarities.py:975-1005)."""
    pa = pytest.importorskip("pyarrow")
    rng = np.random.default_rng(59)
    words = [w.decode() for w in support.random_strings(rng, 40, 0, 90, b"abcdefgh")]
    engine = szb.LevenshteinDistances()
    want = support.oracle_levenshtein([w.encode() for w in words[4:30]], [w.encode() for w in words[10:40]])
    for kind in (pa.string(), pa.large_string()):
        array = pa.array(words, type=kind)
        got = engine(array.slice(4, 26), array.slice(10, 30), scope)
        assert np.array_equal(got, want)
    utf8 = szb.LevenshteinDistancesUTF8()
    greek = ["αβγ", "αβ", "", "γβα", "abc"]
    assert np.array_equal(utf8(pa.array(greek), None, scope), support.oracle_levenshtein_utf8([g.encode() for g in greek]))


def test_retrieval_shapes_trade_roles(scope):
    """A handful of queries against many candidates: with symmetric costs the engine keeps the queries resident and
    streams the candidates (results transposed on the way home); asymmetric tables and row shards keep their roles."""
    rng = np.random.default_rng(61)
    byte_to_class, blosum = scoring.blosum62()
    skewed = blosum.copy()
    skewed[3, 7] += 2  # no longer symmetric: roles must stay
    for n_queries, n_candidates, low, high in ((1, 700, 0, 300), (3, 257, 100, 1200), (31, 40, 0, 90), (2, 3, 2000, 5000)):
        dna_q = support.random_strings(rng, n_queries, low, high, support.DNA)
        dna_c = support.mutated_copies(rng, (dna_q * (n_candidates // n_queries + 1))[:n_candidates], 0.2, support.DNA)
        for costs, kwargs in (((0, 1, 1, 1), {}), ((0, 1, 1, 1), {"bit_parallel": False}), ((1, 3, 3, 3), {}), ((1, 3, 4, 2), {})):
            out = np.full((n_queries, n_candidates + 5), 77, dtype=np.uint64)  # a wider matrix: the stride is honoured
            szb.LevenshteinDistances(*costs, **kwargs)(dna_q, dna_c, scope, out=out[:, :n_candidates])
            assert np.array_equal(out[:, :n_candidates], support.oracle_levenshtein(dna_q, dna_c, *costs)), (n_queries, costs)
            assert (out[:, n_candidates:] == 77).all()
        prot_q = support.random_strings(rng, n_queries, min(low, 1000), min(high, 1500), support.PROTEIN)
        prot_c = support.random_strings(rng, n_candidates, min(low, 1000), min(high, 1500), support.PROTEIN)
        for table in (blosum, skewed):
            for engine_type, local in ((szb.NeedlemanWunschScores, False), (szb.SmithWatermanScores, True)):
                for gaps in ((-4, -4), (-4, -1)):
                    got = engine_type(byte_to_class, table, open=gaps[0], extend=gaps[1])(prot_q, prot_c, scope)
                    want = support.oracle_weighted(prot_q, prot_c, byte_to_class, table, gaps[0], gaps[1], local)
                    assert np.array_equal(got, want), (n_queries, local, gaps, table is skewed)
    greek = ["αβγδ" * 20, "δγβα" * 7]
    many = ["".join(rng.choice(list("αβγδε"), size=int(n))) for n in rng.integers(0, 120, size=90)]
    got = szb.LevenshteinDistancesUTF8()(greek, many, scope)
    assert np.array_equal(got, support.oracle_levenshtein_utf8([g.encode() for g in greek], [m.encode() for m in many]))
    # A row shard of a retrieval-shaped call keeps the ordinary roles and fills only its rows.
    queries = support.random_strings(rng, 4, 50, 200, support.DNA)
    candidates = support.random_strings(rng, 300, 50, 200, support.DNA)
    out = np.zeros((4, 300), dtype=np.uint64)
    engine = szb.LevenshteinDistances()
    engine.score_rows(queries, candidates, 1, 3, device=scope, out=out)
    want = support.oracle_levenshtein(queries, candidates)
    assert np.array_equal(out[1:3], want[1:3]) and (out[0] == 0).all() and (out[3] == 0).all()


def test_row_shards_reassemble_the_matrix(scope):
    rng = np.random.default_rng(23)
    queries = support.random_strings(rng, 13, 0, 200, support.DNA)
    candidates = support.random_strings(rng, 7, 0, 200, support.DNA)
    engine = szb.LevenshteinDistances(1, 3, 3, 3)
    out = np.zeros((13, 7), dtype=np.uint64)
    for begin, end in ((0, 4), (4, 5), (5, 13)):
        engine.score_rows(queries, candidates, begin, end, device=scope, out=out)
    assert np.array_equal(out, support.oracle_levenshtein(queries, candidates, 1, 3, 3, 3))


def test_very_long_and_very_many(scope):
    """Extremes of the scheduler: strings far beyond one pass of any shape (tens of thousands of characters: many
    passes, 16-bit -> 32-bit tier switch) and thousands of tiny strings (one job per slot, many empty ones)."""
    rng = np.random.default_rng(47)
    long_q = [support.random_strings(rng, 1, n, n, support.DNA)[0] for n in (12000, 40000, 300)]
    long_c = support.mutated_copies(rng, long_q[:2], 0.05, support.DNA) + [b"ACGT" * 9000]
    for costs, kwargs in (((0, 1, 1, 1), {}), ((0, 1, 1, 1), {"bit_parallel": False}), ((1, 3, 4, 2), {})):
        got = szb.LevenshteinDistances(*costs, **kwargs)(long_q, long_c, device=scope)
        assert np.array_equal(got, support.oracle_levenshtein(long_q, long_c, *costs)), (costs, kwargs)
    byte_to_class, table = scoring.blosum62()
    prot_q = [support.random_strings(rng, 1, n, n, support.PROTEIN)[0] for n in (9000, 150)]
    prot_c = support.mutated_copies(rng, prot_q[:1], 0.1, support.PROTEIN) + support.random_strings(rng, 1, 5000, 5000, support.PROTEIN)
    for engine_type, local in ((szb.NeedlemanWunschScores, False), (szb.SmithWatermanScores, True)):
        got = engine_type(byte_to_class, table, open=-4, extend=-1)(prot_q, prot_c, device=scope)
        assert np.array_equal(got, support.oracle_weighted(prot_q, prot_c, byte_to_class, table, -4, -1, local))
    tiny_q = support.random_strings(rng, 3000, 0, 9, support.DNA)
    tiny_c = support.random_strings(rng, 2500, 0, 9, support.DNA)
    assert np.array_equal(szb.LevenshteinDistances()(tiny_q, tiny_c, device=scope), support.oracle_levenshtein(tiny_q, tiny_c))
    got = szb.NeedlemanWunschScores(byte_to_class, table, open=-4, extend=-1)(tiny_q[:700], tiny_c[:900], device=scope)
    assert np.array_equal(got, support.oracle_weighted(tiny_q[:700], tiny_c[:900], byte_to_class, table, -4, -1, False))


def test_engine_follows_the_scope_across_devices():
    """One engine, scopes on two GPUs (legal per the reference's threading contract, types.cuh:433-444): the engine's
    grow-only buffers must move with the scope."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(43)
    queries = support.random_strings(rng, 9, 0, 300, support.DNA)
    candidates = support.random_strings(rng, 8, 0, 300, support.DNA)
    want = support.oracle_levenshtein(queries, candidates, 1, 3, 3, 3)
    engine = szb.LevenshteinDistances(1, 3, 3, 3)
    for device in (0, 1, 0):
        assert np.array_equal(engine(queries, candidates, device=szb.DeviceScope(gpu_device=device)), want)


def test_full_size_tile_properties_and_sampled_parity(scope):
    """BASELINE config 2 scale (q316 x c317, |s| ~ 1024): size-independent properties over the whole matrix plus
    exact parity on sampled rows (the oracle would need minutes for all 1e11 cells)."""
    rng = np.random.default_rng(42)
    queries = support.random_strings(rng, 316, 960, 1088, support.DNA)
    candidates = support.mutated_copies(rng, queries[:200], 0.1, support.DNA) + support.random_strings(
        rng, 117, 960, 1088, support.DNA)
    engine = szb.LevenshteinDistances()
    matrix = engine(queries, candidates, device=scope)
    lengths_q = np.array([len(s) for s in queries])[:, None]
    lengths_c = np.array([len(s) for s in candidates])[None, :]
    assert (matrix >= np.abs(lengths_q - lengths_c)).all() and (matrix <= np.maximum(lengths_q, lengths_c)).all()
    rows = [0, 57, 199, 315]
    want = support.oracle_levenshtein([queries[r] for r in rows], candidates)
    assert np.array_equal(matrix[rows], want)
    # Swapping the sides transposes the matrix (unit costs are symmetric).
    transposed = engine(candidates, queries, device=scope)
    assert np.array_equal(transposed, matrix.T)
    stats = engine.last_call_stats
    assert stats.cells == int(lengths_q.sum()) * int(lengths_c.sum()) and stats.kernel_launches >= 1


def test_large_pageable_corpus_is_staged_in_parallel(scope):
    """A candidate blob of more than 4 MB in pageable memory takes the multi-threaded pinned staging path (a retrieval
    corpus against a couple of queries, which also trades roles); results must not depend on how the bytes travelled."""
    rng = np.random.default_rng(67)
    corpus = support.random_strings(rng, 4600, 900, 1000, support.DNA)
    queries = support.random_strings(rng, 2, 400, 600, support.DNA)
    tape = szb.Tape(corpus)
    assert tape.data.nbytes > 4 << 20
    engine = szb.LevenshteinDistances()
    got = engine(queries, tape, device=scope)
    assert np.array_equal(got, support.oracle_levenshtein(queries, tape))
    assert engine.last_call_stats.h2d_bytes >= tape.data.nbytes
    # The same corpus as the query side of a weighted engine (no role swap: 3 candidates, 4600 rows).
    weird = szb.LevenshteinDistances(1, 3, 3, 3)
    few = support.random_strings(rng, 3, 100, 200, support.DNA)
    assert np.array_equal(weird(tape, few, device=scope), support.oracle_levenshtein(tape, few, 1, 3, 3, 3))


# One engine per recurrence family of kernels.cuh, with the widest register strip the scheduler allows it (launch.cuh).
def _family_engines():
    byte_to_class, blosum = scoring.blosum62()
    binary_classes, heavy = scoring.uniform(b"ab", 100, -100)
    text = bytes(range(33, 113))  # 80 distinct bytes: byte-compare recurrences
    weighted = lambda table_classes, table, local: (
        lambda open, extend: (lambda q, c: support.oracle_weighted(q, c, table_classes, table, open, extend, local)))
    return [
        ("lin_shift", szb.LevenshteinDistances(1, 3, 3, 3), support.DNA, 40,
         lambda q, c: support.oracle_levenshtein(q, c, 1, 3, 3, 3)),
        ("aff_shift", szb.LevenshteinDistances(1, 3, 4, 2), support.DNA, 36,
         lambda q, c: support.oracle_levenshtein(q, c, 1, 3, 4, 2)),
        ("sw_lin", szb.SmithWatermanScores(byte_to_class, blosum, open=-4, extend=-4), support.PROTEIN, 40,
         weighted(byte_to_class, blosum, True)(-4, -4)),
        ("sw_aff", szb.SmithWatermanScores(byte_to_class, blosum, open=-4, extend=-1), support.PROTEIN, 36,
         weighted(byte_to_class, blosum, True)(-4, -1)),
        ("s32_global_lin", szb.NeedlemanWunschScores(binary_classes, heavy, open=-127, extend=-127), b"ab", 32,
         weighted(binary_classes, heavy, False)(-127, -127)),
        ("s32_global_aff", szb.NeedlemanWunschScores(byte_to_class, blosum, open=-2, extend=1), support.PROTEIN, 20,
         weighted(byte_to_class, blosum, False)(-2, 1)),
        ("s32_local_lin", szb.SmithWatermanScores(byte_to_class, blosum, open=1, extend=1), support.PROTEIN, 32,
         weighted(byte_to_class, blosum, True)(1, 1)),
        ("s32_local_aff", szb.SmithWatermanScores(byte_to_class, blosum, open=-2, extend=1), support.PROTEIN, 20,
         weighted(byte_to_class, blosum, True)(-2, 1)),
        ("lin_compare", szb.LevenshteinDistances(1, 3, 3, 3), text, 32,
         lambda q, c: support.oracle_levenshtein(q, c, 1, 3, 3, 3)),
    ]


@pytest.mark.parametrize("family", range(9), ids=[entry[0] for entry in _family_engines()])
def test_every_strip_width_of_every_family(family, scope):
    """Every instantiation `wavefront_kernel<family, K>` the scheduler can pick. A strip of K columns is chosen for the
    candidates whose length K x G covers most tightly, and a call too small to fill the device takes G = 32 lanes whatever
    the length — so each strip width gets one candidate just under K x 32 and one just under K x 8 (one of the two lands on
    it, depending on the device's size), in separate calls so that no launch is folded into a wider one."""
    name, engine, alphabet, widest, oracle = _family_engines()[family]
    rng = np.random.default_rng(700 + family)
    queries = support.random_strings(rng, 5, 150, 330, alphabet) + [b"", alphabet[:1]]
    for strip in range(8, widest + 1, 4):
        for lanes in (32, 8):
            top = strip * lanes
            candidates = [support.random_strings(rng, 1, n, n, alphabet)[0] for n in (top, top - 1, top - 3, max(top - 4 * lanes + 1, 1))]
            candidates += support.mutated_copies(rng, [queries[0]], 0.2, alphabet)  # something similar, so scores vary
            candidates = [c[:top] for c in candidates]  # nothing longer than the strip under test
            got = engine(queries, candidates, device=scope)
            assert np.array_equal(got, oracle(queries, candidates)), (name, strip, lanes)


@pytest.mark.parametrize("table", ["per_warp", "per_cta"])
def test_every_word_count_of_the_bit_parallel_kernels(table, scope):
    """Every instantiation `myers_kernel<W>` / `myers_shared_kernel<W>`, W = 1 .. 10 words per lane: launches of one call are
    folded into the widest one, so each W gets calls of its own whose longest candidate needs exactly W words — at 32 lanes
    (a small call on a large device) and at 4 lanes (the tightest shape)."""
    alphabet = support.DNA if table == "per_warp" else bytes(range(33, 113))
    rng = np.random.default_rng(800 + len(alphabet))
    queries = support.random_strings(rng, 9, 0, 700, alphabet) + [b""]
    engine = szb.LevenshteinDistances()
    for words in range(1, 11):
        for lanes in (32, 4):
            top = 32 * words * lanes
            candidates = [support.random_strings(rng, 1, n, n, alphabet)[0] for n in (top, top - 1, top - 31, top - 33)]
            candidates += [queries[1][:top], b""]
            got = engine(queries, candidates, device=scope)
            assert np.array_equal(got, support.oracle_levenshtein(queries, candidates)), (table, words, lanes)
            assert engine.last_call_stats.pairs_bit_parallel > 0


def test_malformed_calls_are_refused_and_leave_the_engine_usable(scope):
    """Status codes of the reference's contract (stringzillas.h / types.h:810-833) for inputs the library must not score:
    decreasing or missing offsets, a missing result matrix, a row stride below the candidate count, a handle of another
    family, pairs too long for 32-bit cells. After every refusal the same engine scores a valid call."""
    rng = np.random.default_rng(61)
    queries = support.random_strings(rng, 4, 1, 40, support.DNA)
    candidates = support.random_strings(rng, 3, 1, 40, support.DNA)
    want = support.oracle_levenshtein(queries, candidates, 1, 3, 3, 3)
    engine = szb.LevenshteinDistances(1, 3, 3, 3)
    lib = _abi.bind()
    q, c = szb.Tape(queries), szb.Tape(candidates)
    message = ct.c_char_p()
    out = np.zeros((4, 3), dtype=np.uint64)

    def call(tq, tc, results, stride, function=lib.szs_levenshtein_distances_u64tape, handle=None):
        return function(handle or engine._handle, scope.handle, tq, tc, results, stride, ct.byref(message))

    def still_works():
        out[:] = 0
        assert call(q.pointer(), c.pointer(), out.ctypes.data, 3) == 0 and np.array_equal(out, want)

    still_works()
    backwards = q.offsets.copy()
    backwards[1] = backwards[2] + 1  # the third offset now lies before the second
    broken = _abi.SequenceU64Tape(q.data.ctypes.data, backwards.ctypes.data, len(q))
    assert call(ct.byref(broken), c.pointer(), out.ctypes.data, 3) == -15 and b"decrease" in message.value
    still_works()
    missing = _abi.SequenceU64Tape(q.data.ctypes.data, None, len(q))
    assert call(ct.byref(missing), c.pointer(), out.ctypes.data, 3) == -15 and b"null" in message.value
    still_works()
    assert call(q.pointer(), c.pointer(), None, 3) == -15 and b"null" in message.value
    still_works()
    assert call(q.pointer(), c.pointer(), out.ctypes.data, 2) == -15 and b"stride" in message.value
    still_works()
    # A Levenshtein handle passed to the Needleman-Wunsch entry point.
    assert call(q.pointer(), c.pointer(), out.ctypes.data, 3, function=lib.szs_needleman_wunsch_scores_u64tape) != 0
    assert b"another algorithm" in message.value
    still_works()
    # (m + n + 4) x the largest cost magnitude must stay below 2^27 for the 32-bit tier: refused before anything runs.
    heavy = szb.LevenshteinDistances(0, 127, 127, 127)
    long_q, long_c = szb.Tape([b"a" * 600_000]), szb.Tape([b"b" * 500_000])
    one = np.zeros((1, 1), dtype=np.uint64)
    assert call(long_q.pointer(), long_c.pointer(), one.ctypes.data, 1, handle=heavy._handle) == -14  # sz_overflow_risk_k
    assert b"32-bit" in message.value
    assert call(q.pointer(), c.pointer(), out.ctypes.data, 3, handle=heavy._handle) == 0
    assert np.array_equal(out, support.oracle_levenshtein(queries, candidates, 0, 127, 127, 127))
    # Empty sides are not errors: nothing is written, nothing is read.
    empty = szb.Tape([])
    assert call(empty.pointer(), c.pointer(), None, 3) == 0 and call(q.pointer(), empty.pointer(), None, 0) == 0


def test_inputs_too_large_for_the_host_side_shortcuts(scope):
    """More text than the host scans or packs itself (256 KB packed arena, 512 KB alphabet scan): the bytes are uploaded as
    blobs, presence flags come from the device, UTF-8 engines decode there; callback sequences are gathered through the
    engine's pinned buffer. Many short strings keep the cross products small."""
    rng = np.random.default_rng(67)
    alphabet = list("abcdefgh абвгдежз αβγδεζηθ")
    def text(n):
        return "".join(alphabet[i] for i in rng.integers(0, len(alphabet), size=n)).encode()
    candidates = [text(int(n)) for n in rng.integers(20, 90, size=7000)]  # ~600 KB of UTF-8
    queries = [text(60), text(35), b"plain ascii", b""]
    assert sum(map(len, candidates)) > 520 * 1024
    for costs in ((0, 1, 1, 1), (1, 3, 3, 3)):
        got = szb.LevenshteinDistancesUTF8(*costs)(queries, candidates, device=scope)
        assert np.array_equal(got, support.oracle_levenshtein_utf8(queries, candidates, *costs)), costs
        got = szb.LevenshteinDistancesUTF8(*costs)(candidates[:3000], queries, device=scope)  # the large side as queries
        assert np.array_equal(got, support.oracle_levenshtein_utf8(candidates[:3000], queries, *costs)), costs
    byte_level = szb.LevenshteinDistances(1, 3, 3, 3)
    assert np.array_equal(byte_level(queries, candidates, device=scope), support.oracle_levenshtein(queries, candidates, 1, 3, 3, 3))
    # The same volume through callback sequences (one callback pair per string).
    lib = _abi.bind()
    buffers = [ct.create_string_buffer(s, len(s)) if s else ct.create_string_buffer(1) for s in candidates]
    start = _abi.SEQUENCE_START(lambda _, i: ct.addressof(buffers[i]))
    length = _abi.SEQUENCE_LENGTH(lambda _, i: len(candidates[i]))
    sequence = _abi.Sequence(None, len(candidates), start, length)
    few = [ct.create_string_buffer(s, len(s)) if s else ct.create_string_buffer(1) for s in queries]
    few_start = _abi.SEQUENCE_START(lambda _, i: ct.addressof(few[i]))
    few_length = _abi.SEQUENCE_LENGTH(lambda _, i: len(queries[i]))
    few_sequence = _abi.Sequence(None, len(queries), few_start, few_length)
    got = np.zeros((len(queries), len(candidates)), dtype=np.uint64)
    message = ct.c_char_p()
    status = lib.szs_levenshtein_distances(byte_level._handle, scope.handle, ct.byref(few_sequence), ct.byref(sequence),
                                           got.ctypes.data, len(candidates), ct.byref(message))
    assert status == 0 and np.array_equal(got, support.oracle_levenshtein(queries, candidates, 1, 3, 3, 3))



def test_alphabet_detection_sees_late_and_rare_bytes(scope):
    """The host scans small inputs for the byte values they contain a block at a time (host_scan.cpp): a value that turns up
    once, at the very end, in the middle of a block or in a tail shorter than a block must still get a class of its own."""
    rng = np.random.default_rng(71)
    queries = support.random_strings(rng, 6, 20, 60, b"ACGTN")
    candidates = [b"A" * 70_000 + b"C", b"A" * 999 + b"G" + b"A" * 300, b"T" * 255 + b"N", b"A" * 256, b"AC" * 128 + b"G",
                  b"N", b"", bytes(range(65, 75)) * 30 + b"z"]
    for costs in ((0, 1, 1, 1), (1, 3, 3, 3), (1, 3, 4, 2)):
        got = szb.LevenshteinDistances(*costs)(queries, candidates, device=scope)
        assert np.array_equal(got, support.oracle_levenshtein(queries, candidates, *costs)), costs
        symmetric = szb.LevenshteinDistances(*costs)(candidates[1:], device=scope)
        assert np.array_equal(symmetric, support.oracle_levenshtein(candidates[1:], None, *costs)), costs


@pytest.mark.parametrize("seed", range(10))
def test_randomized_runes_and_row_shards(seed, scope):
    """Seeded random configurations of what `test_randomized_differential` leaves out: the UTF-8 engine over mixed scripts
    (ASCII only, a few dozen runes, hundreds of runes — three different representations inside the library), and matrices
    assembled from row shards (`score_rows` over random cuts, symmetric calls included; the threaded device group)."""
    rng = np.random.default_rng(3000 + seed)
    scripts = ["abcdefghij ", "abcdefgh абвгдежз αβγδεζηθ 🙂🚀", "".join(chr(0x4E00 + i) for i in range(400)) + "abc"]
    for _ in range(3):
        script = scripts[int(rng.integers(0, len(scripts)))]
        top = int(rng.choice([6, 40, 200, 700]))
        text = lambda: "".join(script[i] for i in rng.integers(0, len(script), size=int(rng.integers(0, top + 1)))).encode()
        queries = [text() for _ in range(int(rng.integers(1, 12)))]
        candidates = None if rng.random() < 0.3 else [text() for _ in range(int(rng.integers(1, 12)))]
        costs = (0, 1, 1, 1) if rng.random() < 0.5 else (int(rng.integers(0, 2)), int(rng.integers(1, 6)), int(rng.integers(1, 6)), int(rng.integers(1, 6)))
        got = szb.LevenshteinDistancesUTF8(*costs)(queries, candidates, device=scope)
        assert np.array_equal(got, support.oracle_levenshtein_utf8(queries, candidates, *costs)), (script[:8], top, costs)
    # Row shards: random cuts of the query rows, each shard scored on its own, must tile the whole matrix.
    byte_to_class, table = scoring.blosum62()
    skewed = table.copy()
    skewed[2, 9] -= 3  # asymmetric: a symmetric call must keep the reference's (query, candidate) orientation
    engines = [(szb.LevenshteinDistances(), support.DNA, lambda q, c: support.oracle_levenshtein(q, c)),
               (szb.LevenshteinDistances(1, 3, 4, 2), support.DNA, lambda q, c: support.oracle_levenshtein(q, c, 1, 3, 4, 2)),
               (szb.NeedlemanWunschScores(byte_to_class, skewed, open=-4, extend=-1), support.PROTEIN,
                lambda q, c: support.oracle_weighted(q, c, byte_to_class, skewed, -4, -1, False)),
               (szb.SmithWatermanScores(byte_to_class, table, open=-3, extend=-3), support.PROTEIN,
                lambda q, c: support.oracle_weighted(q, c, byte_to_class, table, -3, -3, True))]
    engine, alphabet, oracle = engines[seed % len(engines)]
    queries = support.random_strings(rng, int(rng.integers(2, 40)), 0, 260, alphabet)
    for candidates in (support.random_strings(rng, int(rng.integers(1, 20)), 0, 260, alphabet), None):
        columns = len(queries) if candidates is None else len(candidates)
        cuts = sorted({0, len(queries), *rng.integers(0, len(queries) + 1, size=3).tolist()})
        out = np.zeros((len(queries), columns), dtype=engine._result_dtype)
        for begin, end in zip(cuts, cuts[1:]):
            engine.score_rows(queries, candidates, begin, end, device=scope, out=out)
        want = oracle(queries, candidates)
        if candidates is None:
            # Shards of a symmetric call fill the cells left of their diagonal block; the caller mirrors.
            lower = np.tril(np.ones_like(want, dtype=bool))
            assert np.array_equal(out[lower], want[lower]), (seed, cuts)
        else:
            assert np.array_equal(out, want), (seed, cuts)
        group = szb.DeviceGroup([0, 0])
        assert np.array_equal(engine.sharded(queries, candidates, group), want), seed
