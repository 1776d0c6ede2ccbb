"""Binding-level drop-in proof: the REFERENCE's own CPython extension — python/stringzillas/*.c
compiled unmodified from /root/reference by oracle/build_ref_python.sh and linked against this repo's
libstringzillas_b200.so — is imported and driven through the reference's Python API on the GPU, with the cases of the
reference's own binding tests (test/similarities.py:205-226 simple cases, 263-293 custom gaps, 368-420 cross-product and
symmetric shapes, 471-640 NW/SW) checked against the CPU oracle. Runs in a child process so that the reference's
`stringzilla` module never shadows anything in the test session."""

from __future__ import annotations

import subprocess
import sys
import sysconfig
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
BINDING_DIR = ROOT / "oracle" / "_ref" / "python"

CHILD = r'''
import sys, time
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import numpy as np
import stringzilla as sz          # the reference's single-string module (provides Strs)
import stringzillas as szs        # the reference's similarity binding, linked against libstringzillas_b200.so
from stringzilla import Strs
from tests import support
from stringzilla_b200 import scoring

scope = szs.DeviceScope(gpu_device=0)
caps = ("cuda",)

# test/similarities.py:205-226 — simple cases through 1x1 calls
engine = szs.LevenshteinDistances(capabilities=caps)
def distance(a, b, engine=engine):
    out = engine(Strs([a]), Strs([b]), device=scope)
    assert out.shape == (1, 1)
    return int(out[0, 0])
assert distance("hello", "hello") == 0 and distance("hello", "hell") == 1 and distance("", "") == 0
assert distance("", "abc") == 3 and distance("abc", "") == 3 and distance("abc", "ac") == 1
assert distance("abc", "a_bc") == 1 and distance("abc", "adc") == 1 and distance("abcdefgABCDEFG", "ABCDEFGabcdefg") == 14

# test/similarities.py:263-293 — custom gaps, closed forms
mismatch, opening, extension = 4, 3, 2
custom = szs.LevenshteinDistances(open=opening, extend=extension, mismatch=mismatch, capabilities=scope)
assert distance("hello", "hell", custom) == opening and distance("", "abc", custom) == opening + 2 * extension
assert distance("abc", "adc", custom) == mismatch
assert distance("abcdefgABCDEFG", "ABCDEFGabcdefg", custom) == min(14 * mismatch, 2 * opening + 12 * extension)

# UTF-8 (the reference skips custom gaps on its CUDA backend; this library supports them, unit costs checked here)
utf8 = szs.LevenshteinDistancesUTF8(capabilities=caps)
assert distance("αβγδ", "αγδ", utf8) == 1 and distance("こんにちは世界", "こんばんは世界", utf8) == 2 and distance("💖", "💗", utf8) == 1

# test/similarities.py:368-420 — cross-product shapes and symmetric self-similarity, against the oracle
rng = np.random.default_rng(42)
for n_q, n_c in ((1, 1), (1, 6), (5, 1), (3, 4), (8, 8), (40, 33)):
    queries = [s.decode() for s in support.random_strings(rng, n_q, 4, 300, b"abcdefghij")]
    candidates = [s.decode() for s in support.random_strings(rng, n_c, 4, 300, b"abcdefghij")]
    matrix = engine(Strs(queries), Strs(candidates), device=scope)
    assert matrix.shape == (n_q, n_c) and matrix.dtype == np.uint64
    assert np.array_equal(matrix, support.oracle_levenshtein([q.encode() for q in queries], [c.encode() for c in candidates]))
    square = engine(Strs(queries), device=scope)
    assert square.shape == (n_q, n_q) and np.array_equal(square, square.T) and not np.diagonal(square).any()
    assert np.array_equal(square, support.oracle_levenshtein([q.encode() for q in queries]))

# NW / SW with BLOSUM62 (test/similarities.py:471-640 compares with the `affine_gaps` package; the oracle here)
byte_to_class, table = scoring.blosum62()
proteins_q = [s.decode() for s in support.random_strings(rng, 9, 0, 400, support.PROTEIN)]
proteins_c = [s.decode() for s in support.random_strings(rng, 7, 0, 400, support.PROTEIN)]
for engine_type, local in ((szs.NeedlemanWunschScores, False), (szs.SmithWatermanScores, True)):
    for gaps in ((-4, -4), (-4, -1)):
        scorer = engine_type(byte_to_class=byte_to_class, class_substitution_costs=table, open=gaps[0], extend=gaps[1], capabilities=caps)
        got = scorer(Strs(proteins_q), Strs(proteins_c), device=scope)
        want = support.oracle_weighted([p.encode() for p in proteins_q], [p.encode() for p in proteins_c], byte_to_class, table, gaps[0], gaps[1], local)
        assert got.dtype == np.int64 and np.array_equal(got, want), (engine_type.__name__, gaps)

if len(sys.argv) > 3 and sys.argv[3] == "quick":  # the emulated library of the CPU tier: no c2-sized call
    print("BINDING_OK reference CPython binding, shape and parity cases only")
    sys.exit(0)

# A managed-memory call the way the binding makes it: `Strs` swapped to the unified allocator (similarities.c:267-271)
dna = Strs([s.decode() for s in support.random_strings(rng, 316, 960, 1088, b"ACGT")])
other = Strs([s.decode() for s in support.random_strings(rng, 317, 960, 1088, b"ACGT")])
engine(dna, other, device=scope)
start = time.perf_counter()
for _ in range(5):
    out = engine(dna, other, device=scope)
seconds = (time.perf_counter() - start) / 5
cells = sum(len(s) for s in dna) * sum(len(s) for s in other)
print(f"BINDING_OK reference CPython binding over libstringzillas_b200.so: q316 x c317 DNA in {seconds * 1e3:.2f} ms = {cells / seconds / 1e9:.0f} GCUPS")
'''


@pytest.mark.gpu
def test_reference_cpython_binding_over_this_library(tmp_path):
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    if not (BINDING_DIR / f"stringzillas{suffix}").exists():
        pytest.skip("oracle/_ref/python is not built (needs /root/reference at build time)")
    script = tmp_path / "child.py"
    script.write_text(CHILD)
    done = subprocess.run([sys.executable, str(script), str(ROOT), str(BINDING_DIR)], capture_output=True, text=True, timeout=600)
    assert done.returncode == 0 and "BINDING_OK" in done.stdout, done.stdout[-2000:] + done.stderr[-4000:]
    print(done.stdout.strip().splitlines()[-1])


def test_reference_cpython_binding_over_the_emulated_library(tmp_path):
    """The same binding on the CPU test tier: the dynamic loader is pointed (LD_LIBRARY_PATH beats the binding's RUNPATH) at a
    directory where `libstringzillas_b200.so` is the emulated build of tests/emulator — the reference's unmodified CPython
    extension then drives the unchanged host driver and kernels under the SIMT emulator. Test infrastructure only."""
    import os

    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    if not (BINDING_DIR / f"stringzillas{suffix}").exists():
        pytest.skip("oracle/_ref/python is not built (needs /root/reference at build time)")
    from tests.emulator import build as emulator_build

    emulated = emulator_build.build()
    (tmp_path / "lib").mkdir()
    (tmp_path / "lib" / "libstringzillas_b200.so").symlink_to(emulated)
    script = tmp_path / "child.py"
    script.write_text(CHILD)
    environment = dict(os.environ, LD_LIBRARY_PATH=str(tmp_path / "lib"), SZB_EMULATED_SMS="2")
    done = subprocess.run([sys.executable, str(script), str(ROOT), str(BINDING_DIR), "quick"], capture_output=True, text=True,
                          timeout=900, env=environment)
    assert done.returncode == 0 and "BINDING_OK" in done.stdout, done.stdout[-2000:] + done.stderr[-4000:]

