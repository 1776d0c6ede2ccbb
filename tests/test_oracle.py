"""CPU suite, part 1: the oracle (oracle/sz_oracle.c) is pinned against the reference's own vectors.

- tests/golden/reference_vectors.json holds matrices produced by the UNMODIFIED reference CPU library
  (tests/golden/make_golden.py), with its serial and Ice Lake tiers agreeing;
- the 15 known-answer Levenshtein vectors come from reference test/similarities.cuh:609-625;
- closed forms follow reference test/similarities.cuh:1042-1058 and test/similarities.py:282-293.
"""

import json

import numpy as np
import pytest

from tests import support

GOLDEN = json.loads((support.GOLDEN_DIR / "reference_vectors.json").read_text())


def oracle_for_case(case):
    queries = [bytes.fromhex(q) for q in case["queries"]]
    candidates = None if case["candidates"] is None else [bytes.fromhex(c) for c in case["candidates"]]
    params = case["params"]
    if case["kind"] == "levenshtein":
        return support.oracle_levenshtein(queries, candidates, *params["costs"]).astype(np.int64)
    if case["kind"] == "levenshtein_utf8":
        return support.oracle_levenshtein_utf8(queries, candidates, *params["costs"]).astype(np.int64)
    return support.oracle_weighted(
        queries, candidates, np.array(params["byte_to_class"], dtype=np.uint8),
        np.array(params["costs"], dtype=np.int8), params["open"], params["extend"],
        case["kind"] == "smith_waterman")  # fmt: skip


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=[c["name"] for c in GOLDEN["cases"]])
def test_oracle_matches_reference_golden_vectors(case):
    assert np.array_equal(oracle_for_case(case), np.array(case["matrix"], dtype=np.int64))


def test_oracle_known_answer_levenshtein_vectors():
    for first, second, distance in GOLDEN["known_levenshtein"]:
        got = support.oracle_levenshtein([first.encode()], [second.encode()])
        assert int(got[0, 0]) == distance, (first, second)


def test_oracle_closed_forms():
    sample = b"ADVERSARIAL"
    for gap in (1, 3, 7):
        assert int(support.oracle_levenshtein([sample], [b""], 0, 1, gap, gap)[0, 0]) == len(sample) * gap
        assert int(support.oracle_levenshtein([sample], [sample], 0, 1, gap, gap)[0, 0]) == 0
    # Custom affine gaps, reference test/similarities.py:282-293.
    mismatch, opening, extension = 4, 3, 2
    distance = lambda a, b: int(support.oracle_levenshtein([a], [b], 0, mismatch, opening, extension)[0, 0])
    assert distance(b"hello", b"hell") == opening
    assert distance(b"", b"abc") == opening + 2 * extension
    assert distance(b"abc", b"adc") == mismatch
    assert distance(b"abcdefgABCDEFG", b"ABCDEFGabcdefg") == min(14 * mismatch, 2 * opening + 12 * extension)


@pytest.mark.skipif(not support.have_reference(), reason="oracle/_ref not built (the reference tree is absent here)")
@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_matches_live_reference_on_random_batches(seed):
    import stringzilla_b200 as szb
    from stringzilla_b200 import scoring

    rng = np.random.default_rng(seed)
    ref = support.REFERENCE_LIBRARY
    scope = szb.DeviceScope(library_path=ref)
    queries = support.random_strings(rng, 12, 0, 400, support.DNA)
    candidates = support.mutated_copies(rng, queries[:9], 0.2, support.DNA)
    for costs in ([0, 1, 1, 1], [1, 3, 3, 3], [1, 3, 4, 2], [0, 7, 2, 2]):
        got = szb.LevenshteinDistances(*costs, library_path=ref)(queries, candidates, device=scope)
        assert np.array_equal(got, support.oracle_levenshtein(queries, candidates, *costs)), costs
    alphabet = list("abcdefgh абвгдежз αβγδεζηθ 数据科学机器 🙂🌍🚀💖")
    text = lambda n: "".join(alphabet[i] for i in rng.integers(0, len(alphabet), size=n)).encode()
    utf8_q = [text(int(n)) for n in rng.integers(0, 200, size=8)]
    utf8_c = [text(int(n)) for n in rng.integers(0, 300, size=7)]
    for costs in ([0, 1, 1, 1], [1, 3, 3, 3], [0, 4, 3, 2]):
        got = szb.LevenshteinDistancesUTF8(*costs, library_path=ref)(utf8_q, utf8_c, device=scope)
        assert np.array_equal(got, support.oracle_levenshtein_utf8(utf8_q, utf8_c, *costs)), costs
    byte_to_class, table = scoring.blosum62()
    proteins = support.random_strings(rng, 10, 0, 300, support.PROTEIN)
    others = support.mutated_copies(rng, proteins[:7], 0.3, support.PROTEIN)
    for engine_type, local in ((szb.NeedlemanWunschScores, False), (szb.SmithWatermanScores, True)):
        for open_, extend in ((-4, -4), (-4, -1), (-11, -1)):
            got = engine_type(byte_to_class, table, open=open_, extend=extend, library_path=ref)(proteins, others, device=scope)
            want = support.oracle_weighted(proteins, others, byte_to_class, table, open_, extend, local)
            assert np.array_equal(got, want), (engine_type.__name__, open_, extend)
