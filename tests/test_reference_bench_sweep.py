"""bench/similarities_abi: the reference bench's measurement (shapes, cell accounting, baseline equality check) one level
up, over the C ABI. On CPU both arms are the reference's own library (tier mask vs tier mask), which exercises every
code path of the harness except the device-timer line; the GPU test puts this repo's library on the accelerated arm."""

from __future__ import annotations

import os
import re
import subprocess
from pathlib import Path

import pytest

from tests import support

ROOT = Path(__file__).resolve().parents[1]
BINARY = ROOT / "bench" / "similarities_abi"
REFERENCE = ROOT / "oracle" / "_ref" / "libstringzillas_ref.so"


def run_benchmark(**overrides) -> str:
    subprocess.run(["make", "-C", str(ROOT / "bench")], check=True, capture_output=True)
    env = dict(os.environ, STRINGWARS_DURATION="0.02", STRINGWARS_STRESS="1", STRINGWARS_STRESS_DURATION="0",
               STRINGWARS_BATCH_PER_CORE="1", STRINGWARS_SYNTHETIC_LENGTH="64", SZS_CPU_CORES="4")
    env.update(overrides)
    done = subprocess.run([str(BINARY)], env=env, capture_output=True, text=True, timeout=300)
    assert done.returncode == 0, done.stdout[-2000:] + done.stderr[-2000:]
    return done.stdout


@pytest.mark.skipif(not support.have_reference(), reason="oracle/_ref is not built")
def test_reference_library_on_both_arms():
    out = run_benchmark(SZS_ACCELERATED_LIB=str(REFERENCE), SZS_BASELINE_LIB=str(REFERENCE))
    assert "All benchmarks finished." in out and "Mismatch" not in out and "> Errors" not in out
    names = re.findall(r"Benchmarking `([^`]+)`", out)
    # 2 cost schemes x 3 Levenshtein variants + 4 weighted variants, each over (cross, diagonal) x 4 batch sizes, x 2 arms
    assert len(names) == (2 * 3 + 4) * 8 * 2
    for expected in ("levenshtein_serial_unit", "affine_levenshtein_cpu_weird", "levenshtein_utf8_cpu_unit",
                     "needleman_wunsch_cpu", "affine_smith_waterman_serial"):
        assert any(name.startswith(expected + ":") for name in names), expected
    assert out.count("> Stress: 1 tiles compared") == len(names) // 2
    # Cell accounting: q x c tokens of 60..68 bytes each => the efficiency line must be present for every benchmark
    assert out.count("> Efficiency:") == len(names)


@pytest.mark.skipif(not support.have_reference(), reason="oracle/_ref is not built")
def test_filter_and_batch_override():
    out = run_benchmark(SZS_ACCELERATED_LIB=str(REFERENCE), SZS_BASELINE_LIB="none", STRINGWARS_STRESS="0",
                        STRINGWARS_FILTER="^smith_waterman_cpu:q", STRINGWARS_BATCH="3,5")
    assert re.findall(r"Benchmarking `([^`]+)`", out) == ["smith_waterman_cpu:q3xc2", "smith_waterman_cpu:q5xc1"]


def known_answers(tmp_path):
    """A dataset of two distinct protein strings, many copies each: whatever tile the harness samples, every cell is one
    of three pairings, so the extremes it prints must be the oracle's."""
    import numpy as np

    from stringzilla_b200 import scoring

    rng = np.random.default_rng(5)
    first, second = support.random_strings(rng, 2, 150, 180, support.PROTEIN)
    dataset = tmp_path / "two_proteins.txt"
    dataset.write_bytes(b"\n".join([first, second] * 200) + b"\n")
    byte_to_class, costs = scoring.blosum62()
    expected = {}
    for name, (open_, extend, local) in {"needleman_wunsch": (-4, -4, False), "smith_waterman": (-4, -4, True),
                                         "affine_needleman_wunsch": (-4, -1, False),
                                         "affine_smith_waterman": (-4, -1, True)}.items():
        scores = support.oracle_weighted([first, second], [first, second], byte_to_class, costs, open_, extend, local)
        expected[name] = {int(v) for v in scores.ravel()}
    for tag, (match, mismatch, gap, open_, extend) in {"unit": (0, 1, 1, 1, 1), "weird": (1, 3, 3, 4, 2)}.items():
        linear = support.oracle_levenshtein([first, second], [first, second], match, mismatch, gap, gap)
        affine = support.oracle_levenshtein([first, second], [first, second], match, mismatch, open_, extend)
        expected[f"levenshtein_@_{tag}"] = expected[f"levenshtein_utf8_@_{tag}"] = {int(v) for v in linear.ravel()}
        expected[f"affine_levenshtein_@_{tag}"] = {int(v) for v in affine.ravel()}
    return dataset, expected


def check_known_answers(out: str, expected: dict, backend: str):
    blocks = re.findall(r"Benchmarking `([^`:]+):[^`]+`:\n(?:> .*\n)*?> Results of the last tile: min (-?\d+), max (-?\d+)", out)
    seen = set()
    for name, low, high in blocks:
        if f"_{backend}" not in name:
            continue
        key = name.replace(f"_{backend}", "_@" if "levenshtein" in name else "")
        assert {int(low), int(high)} <= expected[key], (name, low, high, expected[key])
        seen.add(key)
    assert seen == set(expected)


@pytest.mark.skipif(not support.have_reference(), reason="oracle/_ref is not built")
def test_known_answers_through_the_harness(tmp_path):
    """Guards the harness itself (argument order of the ABI calls, the substitution table it ships): results of the
    reference library driven by the harness must be the oracle's values for the only three pairings in the dataset."""
    dataset, expected = known_answers(tmp_path)
    out = run_benchmark(SZS_ACCELERATED_LIB=str(REFERENCE), SZS_BASELINE_LIB="none", STRINGWARS_STRESS="0",
                        STRINGWARS_DATASET=str(dataset), STRINGWARS_BATCH="6")
    check_known_answers(out, expected, "cpu")


@pytest.mark.gpu
@pytest.mark.skipif(not support.have_reference(), reason="oracle/_ref is not built")
def test_b200_library_against_the_reference_cpu_baseline(tmp_path):
    """Every engine of the reference bench's sweep, this repo's library on GPU 0 against the reference's CPU library on
    identical tiles (the harness prints each mismatching tile and exits non-zero past the limit), plus the known-answer
    dataset. Sorted after the parity tests on purpose: this file checks the measurement tool, those check the product."""
    out = run_benchmark(STRINGWARS_BATCH_PER_CORE="4", STRINGWARS_SYNTHETIC_LENGTH="200", STRINGWARS_BATCH="1,24")
    assert "on GPU 0" in out and "All benchmarks finished." in out
    assert "Mismatch" not in out and "> Errors" not in out
    assert "levenshtein_b200_unit:q" in out and "> Kernel:" in out
    dataset, expected = known_answers(tmp_path)
    out = run_benchmark(SZS_BASELINE_LIB="none", STRINGWARS_STRESS="0", STRINGWARS_DATASET=str(dataset), STRINGWARS_BATCH="6")
    check_known_answers(out, expected, "b200")


@pytest.mark.skipif(not support.have_reference(), reason="oracle/_ref is not built")
def test_emulated_library_against_the_reference_cpu_baseline(tmp_path):
    """The same sweep on the CPU test tier, with the emulated build of the library (tests/emulator) on the accelerated arm:
    every engine x shape of the harness compares its tiles with the reference CPU library's. Test infrastructure only."""
    from tests.emulator import build as emulator_build

    emulated = str(emulator_build.build())
    out = run_benchmark(SZS_ACCELERATED_LIB=emulated, STRINGWARS_BATCH="1,6", SZB_EMULATED_SMS="2")
    assert "All benchmarks finished." in out and "Mismatch" not in out and "> Errors" not in out
    assert "levenshtein_b200_unit:q" in out and out.count("> Stress: 1 tiles compared") >= 20
    dataset, expected = known_answers(tmp_path)
    out = run_benchmark(SZS_ACCELERATED_LIB=emulated, SZS_BASELINE_LIB="none", STRINGWARS_STRESS="0", STRINGWARS_DATASET=str(dataset),
                        STRINGWARS_BATCH="6", SZB_EMULATED_SMS="2")
    check_known_answers(out, expected, "b200")


@pytest.mark.gpu
def test_reference_benchmark_runs_the_b200_engines(tmp_path):
    """the REFERENCE's own benchmark binary (bench/similarities.cu compiled from /root/reference by
    oracle/build_ref_bench.sh, with this library's engines registered through bench/reference_bench_b200.cuh) runs its
    serial-vs-accelerated equality check (bench/similarities.cuh:410-423) over every `*_b200*` engine."""
    binary = ROOT / "oracle" / "_ref" / "ref_bench_similarities"
    if not binary.exists():
        pytest.skip("oracle/_ref/ref_bench_similarities is not built (needs /root/reference at build time)")
    dataset = tmp_path / "proteins.txt"
    subprocess.run(["python", str(ROOT / "tools" / "make_dataset.py"), str(dataset), "900", "90", "160", "protein"], check=True)
    env = dict(os.environ, STRINGWARS_DATASET=str(dataset), STRINGWARS_TOKENS="lines", STRINGWARS_DURATION="1",
               STRINGWARS_STRESS="1", STRINGWARS_STRESS_DURATION="1", STRINGWARS_STRESS_DIR=str(tmp_path),
               STRINGWARS_FILTER="_b200", STRINGWARS_BATCH="24")
    done = subprocess.run([str(binary)], env=env, capture_output=True, text=True, timeout=420)
    out = done.stdout + done.stderr
    assert done.returncode == 0, out[-3000:]
    assert "All benchmarks finished." in out and "Mismatch" not in out and "> Errors" not in out, out[-3000:]
    names = re.findall(r"Benchmarking (?:\x1b\[1m)?`([^`]+)`", out)  # the reference prints the name in bold
    for expected in ("levenshtein_b200_unit", "affine_levenshtein_b200_weird", "needleman_wunsch_b200", "smith_waterman_b200",
                     "affine_needleman_wunsch_b200", "affine_smith_waterman_b200"):
        assert any(name.startswith(expected + ":") for name in names), (expected, names)
    assert out.count("> Kernel:") >= len([n for n in names if "_b200" in n])  # device time reported through cuda_status_t
