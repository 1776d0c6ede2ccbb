#!/usr/bin/env python3
"""Generates tests/golden/reference_vectors.json by running the UNMODIFIED reference CPU library
(oracle/_ref/libstringzillas_ref.so, built by oracle/build_ref.sh from /root/reference) through its own C ABI.

Run here (the reference tree exists only in the build container):  python tests/golden/make_golden.py
Every case is computed twice — capability mask `serial` and mask `szs_capabilities()` (Ice Lake tier on AVX-512 hosts)
— and the two must agree before the vector is written, as the reference's own tests demand
(test/similarities.cuh:543-545). Inputs are seeded; strings are stored hex-encoded.
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

import stringzilla_b200 as szb  # noqa: E402
from stringzilla_b200 import _abi, scoring  # noqa: E402
from tests import support  # noqa: E402

REF = support.REFERENCE_LIBRARY


def run_reference(kind, queries, candidates, params):
    results = []
    serial_scope = szb.DeviceScope(library_path=REF)
    for caps in (_abi.CAP_SERIAL, szb.capabilities(REF)):
        if kind == "levenshtein":
            engine = szb.LevenshteinDistances(*params["costs"], capabilities=caps, library_path=REF)
        elif kind == "levenshtein_utf8":
            engine = szb.LevenshteinDistancesUTF8(*params["costs"], capabilities=caps, library_path=REF)
        else:
            engine_type = szb.NeedlemanWunschScores if kind == "needleman_wunsch" else szb.SmithWatermanScores
            engine = engine_type(np.array(params["byte_to_class"], dtype=np.uint8),
                                 np.array(params["costs"], dtype=np.int8).reshape(32, 32),
                                 open=params["open"], extend=params["extend"], capabilities=caps, library_path=REF)
        results.append(engine(queries, candidates, device=serial_scope))
    assert np.array_equal(results[0], results[1]), f"reference serial and SIMD tiers disagree for {kind}"
    return results[0]


def main():
    rng = np.random.default_rng(20260922)
    cases = []

    def add(name, kind, queries, candidates, **params):
        matrix = run_reference(kind, queries, candidates, params)
        cases.append({
            "name": name, "kind": kind,
            "queries": [q.hex() for q in queries],
            "candidates": None if candidates is None else [c.hex() for c in candidates],
            "params": params, "matrix": matrix.astype(np.int64).tolist(),
        })
        print(f"{name}: {matrix.shape}")

    blosum_map, blosum = scoring.blosum62()
    blosum_params = dict(byte_to_class=blosum_map.tolist(), costs=blosum.reshape(-1).tolist())

    # Degenerate shapes (reference test/similarities.cuh:1012-1020 exercises the same situations).
    tiny = [b"", b"A", b"ABC", b"ABCABC", b"ABCXBC", b"ADVERSARIAL", b"", b"B"]
    add("lev_unit_degenerate", "levenshtein", tiny, tiny[::-1], costs=[0, 1, 1, 1])
    add("lev_affine_degenerate", "levenshtein", tiny, tiny[::-1], costs=[0, 4, 3, 2])
    add("lev_unit_symmetric", "levenshtein", tiny, None, costs=[0, 1, 1, 1])

    dna_q = support.random_strings(rng, 9, 0, 300, support.DNA)
    dna_c = support.mutated_copies(rng, dna_q[:7], 0.15, support.DNA) + support.random_strings(rng, 4, 1, 90, support.DNA)
    add("lev_unit_dna", "levenshtein", dna_q, dna_c, costs=[0, 1, 1, 1])
    add("lev_weird_linear_dna", "levenshtein", dna_q, dna_c, costs=[1, 3, 3, 3])   # bench/similarities.cuh:464-467
    add("lev_weird_affine_dna", "levenshtein", dna_q, dna_c, costs=[1, 3, 4, 2])   # test/similarities.cuh "weird" affine

    long_q = support.random_strings(rng, 3, 900, 1300, support.DNA)
    long_c = support.mutated_copies(rng, long_q, 0.1, support.DNA) + support.random_strings(rng, 2, 1000, 2500, support.DNA)
    add("lev_unit_long", "levenshtein", long_q, long_c, costs=[0, 1, 1, 1])
    add("lev_affine_long", "levenshtein", long_q, long_c, costs=[0, 2, 5, 1])
    # High-magnitude costs push cells past 16 bits at ~700 characters (test/similarities.cuh:2023-2035).
    wide_q = support.random_strings(rng, 2, 650, 750, b"ab")
    wide_c = support.random_strings(rng, 3, 650, 750, b"ab")
    add("lev_high_cost_wide_cells", "levenshtein", wide_q, wide_c, costs=[0, 100, 100, 100])

    text_q = ["the quick brown fox jumps over the lazy dog".encode(), "Zürich — Łódź — Ærøskøbing".encode(),
              "naïve café, jalapeño & piñata".encode(), "数据科学 123".encode(), b"\x00\x01\x02\xff\xfe binary"]
    text_c = ["the quick brown dog jumps over the lazy fox".encode(), "Zurich - Lodz - Aeroskobing".encode(),
              "naive cafe, jalapeno & pinata".encode(), "數據科學 321".encode(), b"\x00\x02\x01\xfe\xff binary!"]
    add("lev_unit_bytes_of_text", "levenshtein", text_q, text_c, costs=[0, 1, 1, 1])

    # Rune-level distances: accents, combining marks, CJK, emoji, mixed scripts, plus pure ASCII (the byte fallback).
    utf8_q = [t.encode() for t in ("αβγδ", "école", "Schön", "Data科学123", "🙂🌍🚀", "💖", "München", "façade", "こんにちは世界",
                                   "plain ascii text", "", "naïve café, jalapeño & piñata", "Zürich — Łódź — Ærøskøbing",
                                   "Ереван, თბილისი και Αθήνα – три столицы", "数据科学与机器学习" * 9)]
    utf8_c = [t.encode() for t in ("αγδ", "e\u0301cole", "Scho\u0308n", "Data科學321", "🙂🌎✨", "💗", "Muenchen", "facade", "こんばんは世界",
                                   "plain ascii texts", "ß", "naive cafe, jalapeno & pinata", "Zurich - Lodz - Aeroskobing",
                                   "Երևան, თბილისი, και Αθήνα – 3 մայրաքաղաք", "數據科學與機器學習" * 8)]
    add("lev_utf8_unit_mixed_scripts", "levenshtein_utf8", utf8_q, utf8_c, costs=[0, 1, 1, 1])
    add("lev_utf8_affine_mixed_scripts", "levenshtein_utf8", utf8_q, utf8_c, costs=[0, 4, 3, 2])
    add("lev_utf8_weird_linear_symmetric", "levenshtein_utf8", utf8_q, None, costs=[1, 3, 3, 3])
    cyrillic = "абвгдежзийклмнопрстуфхцчшщъыьэюя "
    ru_q = ["".join(cyrillic[i] for i in rng.integers(0, len(cyrillic), size=int(n))).encode() for n in rng.integers(0, 400, size=6)]
    ru_c = support.mutated_copies(rng, [q.decode() for q in ru_q][:0], 0.0, b"") or [
        "".join(cyrillic[i] for i in rng.integers(0, len(cyrillic), size=int(n))).encode() for n in rng.integers(1, 700, size=5)]
    add("lev_utf8_unit_cyrillic_long", "levenshtein_utf8", ru_q, ru_c, costs=[0, 1, 1, 1])
    add("lev_utf8_ascii_only", "levenshtein_utf8", dna_q, dna_c, costs=[0, 1, 1, 1])

    prot_q = support.random_strings(rng, 8, 0, 260, support.PROTEIN)
    prot_c = support.mutated_copies(rng, prot_q[:6], 0.25, support.PROTEIN) + support.random_strings(rng, 3, 1, 120, support.PROTEIN)
    for kind in ("needleman_wunsch", "smith_waterman"):
        add(f"{kind}_blosum62_linear", kind, prot_q, prot_c, open=-4, extend=-4, **blosum_params)
        add(f"{kind}_blosum62_affine", kind, prot_q, prot_c, open=-4, extend=-1, **blosum_params)
        add(f"{kind}_blosum62_symmetric", kind, prot_q, None, open=-4, extend=-1, **blosum_params)
    long_p = support.random_strings(rng, 2, 1100, 1500, support.PROTEIN)
    long_pc = support.mutated_copies(rng, long_p, 0.2, support.PROTEIN) + support.random_strings(rng, 1, 1800, 2100, support.PROTEIN)
    add("needleman_wunsch_blosum62_affine_long", "needleman_wunsch", long_p, long_pc, open=-4, extend=-1, **blosum_params)
    add("smith_waterman_blosum62_affine_long", "smith_waterman", long_p, long_pc, open=-4, extend=-1, **blosum_params)
    # i16 -> i32 edge for BLOSUM62 sits near a combined length of 2975 (test/similarities.cuh:1441-1453).
    edge_p = support.random_strings(rng, 2, 1480, 1500, support.PROTEIN)
    add("needleman_wunsch_cell_width_edge", "needleman_wunsch", edge_p, support.mutated_copies(rng, edge_p, 0.05, support.PROTEIN),
        open=-4, extend=-4, **blosum_params)

    # An asymmetric substitution table: cost[class(query)][class(candidate)] must not be transposed.
    asym_map, asym = scoring.uniform(b"ACGT", 2, -1)
    asym = asym.copy()
    asym[1, 2], asym[2, 1] = -3, 1
    asym[3, 4], asym[4, 3] = 0, -2
    asym_params = dict(byte_to_class=asym_map.tolist(), costs=asym.reshape(-1).tolist())
    add("needleman_wunsch_asymmetric", "needleman_wunsch", dna_q, dna_c, open=-2, extend=-2, **asym_params)
    add("smith_waterman_asymmetric_affine", "smith_waterman", dna_q, dna_c, open=-3, extend=-1, **asym_params)
    add("needleman_wunsch_asymmetric_symmetric_mode", "needleman_wunsch", dna_q, None, open=-2, extend=-1, **asym_params)

    # Known-answer vectors: reference test/similarities.cuh:609-625 (first, second, hand-verified unit distance).
    known = [("", "", 0), ("ABC", "ABC", 0), ("A", "=", 1), ("", "ABC", 3), ("ABC", "", 3), ("ABC", "AABC", 1),
             ("ABC", "ABCC", 1), ("ABC", "AC", 1), ("ABC", "AXBC", 1), ("ABC", "AXC", 1), ("ABCDEFG", "ABCXEFG", 1),
             ("ggbuzgjux{}l", "gbuzgjux{}l", 1), ("APPLE", "APLE", 1), ("LISTEN", "SILENT", 4),
             ("ATCA", "CTACTCACCC", 6)]
    for first, second, distance in known:
        got = run_reference("levenshtein", [first.encode()], [second.encode()], dict(costs=[0, 1, 1, 1]))
        assert int(got[0, 0]) == distance
    out = {"generator": "tests/golden/make_golden.py", "reference": "StringZilla 5.1.2 CPU library (serial and Ice Lake masks agree)",
           "known_levenshtein": [[a, b, d] for a, b, d in known], "cases": cases}
    path = Path(__file__).resolve().parent / "reference_vectors.json"
    path.write_text(json.dumps(out, separators=(",", ":")))
    print(f"wrote {path} ({path.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
