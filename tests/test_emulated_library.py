"""CPU suite, part 6: the WHOLE library on the CPU test tier.

`tests/emulator/` compiles the unchanged sources of `stringzilla_b200/csrc/` — host driver, C ABI, every kernel — with g++
against a lock-step SIMT emulator (one fiber per CUDA thread, warp collectives and CTA barriers as rendezvous points, the
CUDA runtime over host memory; inline PTX lives in one header, `ptx.cuh`, whose emulator counterpart states in C++ what
each instruction computes). The result is TEST INFRASTRUCTURE: about 10^4 times slower than a B200, loaded by nothing in
`stringzilla_b200/`, and of no use as a fallback — what it buys is that kernel LOGIC (recurrence transforms, boundaries,
multi-pass hand-over, match tables, harvest, the scheduler's shapes and jobs, role swaps, mirrors, row shards) is compared
with the oracle and with the reference's golden vectors on every machine, GPU or not.

The parity tests themselves are the GPU suite's, unchanged: they run here in a subprocess whose `SZS_B200_LIBRARY` points
at the emulated build. Device memory comes from a hook of the emulated device
(`support.DeviceBuffer`; torch tensors on a real GPU). Left out: the tests that link against the real library or need two
physical devices, and the ones too heavy for an emulator (a 100k-character pair, full-size tiles, retrieval corpora).
"""

import importlib.util
import os
import re
import subprocess
import sys

import pytest

from tests import support

# Test functions of the GPU suite that run against the emulated library.
SELECTED = [
    "test_golden_vectors_of_the_reference",
    "test_known_answer_levenshtein_vectors",
    "test_levenshtein_random_cross_products",
    "test_bit_parallel_levenshtein_shapes",
    "test_weighted_random_cross_products",
    "test_mixed_lengths_one_call",
    "test_shapes_and_degenerate_inputs",
    "test_symmetric_self_similarity",
    "test_symmetric_triangle_scheduling",
    "test_asymmetric_substitution_table",
    "test_bit_parallel_levenshtein_large_alphabets",
    "test_large_alphabets_and_wide_cells",
    "test_unusual_cost_signs_take_the_literal_tier",
    "test_utf8_levenshtein_rune_level",
    "test_every_input_layout_of_the_abi",
    "test_callback_sequences_outside_host_memory",
    "test_memory_kinds_and_tape_slices",
    "test_inputs_too_large_for_the_host_side_shortcuts",
    "test_randomized_differential",
    "test_arrow_arrays_as_inputs",
    "test_row_shards_reassemble_the_matrix",
    "test_sharded_call_with_three_shards_on_one_device",
    "test_every_strip_width_of_every_family",
    "test_every_word_count_of_the_bit_parallel_kernels",
    "test_malformed_calls_are_refused_and_leave_the_engine_usable",
    "test_alphabet_detection_sees_late_and_rare_bytes",
    "test_randomized_runes_and_row_shards",
]


@pytest.fixture(scope="module")
def emulated_library():
    from tests.emulator import build as emulator_build

    try:
        return emulator_build.build()
    except (RuntimeError, OSError, subprocess.CalledProcessError) as error:  # no g++, or a source the shim cannot follow
        pytest.fail(f"the emulated library did not build: {error}")


def test_emulated_library_exports_the_abi(emulated_library):
    from stringzilla_b200 import _abi

    lib = _abi.bind(emulated_library)
    missing = [name for name in _abi.REFERENCE_SYMBOLS + _abi.B200_SYMBOLS if not hasattr(lib, name)]
    assert not missing, missing
    # The emulated device: one "GPU" of compute capability 10.0.
    assert lib.szs_capabilities_runtime() & _abi.CAP_CUDA


@pytest.mark.parametrize("schedule", ["in_order", "shuffled"])
def test_gpu_parity_suite_on_the_emulated_library(schedule, emulated_library):
    """`shuffled`: warps and lanes take their turns in a pseudo-random, ever-changing order (SZB_EMULATOR_SHUFFLE) on a
    148-SM device (so that small calls take the latency shapes they take on a B200; the in-order pass has 2 SMs). Correctly synchronised kernels compute the same under every order; a kernel that leans on an accident
    of the default order — a missing barrier, a read ahead of a neighbour's write — fails parity here."""
    support.build_oracle()
    expression = " or ".join(f"({name})" for name in SELECTED)
    environment = dict(os.environ, SZS_B200_LIBRARY=str(emulated_library), SZB_EMULATED_SMS="2")
    if schedule == "shuffled":
        environment.update(SZB_EMULATOR_SHUFFLE="20260923", SZB_EMULATED_SMS="148")
    command = [sys.executable, "-m", "pytest", "tests/test_parity_gpu.py", "tests/test_multi_gpu.py", "-m", "gpu", "-q", "-x",
               "-p", "no:cacheprovider", "-k", expression]
    if importlib.util.find_spec("xdist"):  # every worker process loads its own emulator
        command += ["-n", str(max(1, min(4, (os.cpu_count() or 2) // 2)))]
    ran = subprocess.run(command, cwd=support.ROOT, env=environment, capture_output=True, text=True, timeout=1500)
    tail = (ran.stdout + ran.stderr)[-4000:]
    assert ran.returncode == 0, tail
    summary = re.search(r"(\d+) passed", ran.stdout)
    assert summary and int(summary.group(1)) >= 139, tail
    assert "failed" not in ran.stdout and "error" not in ran.stdout.lower(), tail


def test_multi_rank_exchange_with_ranks_as_threads(emulated_library):
    """`szs_b200_comm_scores_u64tape` for worlds of 2 and 3 ranks on the CPU tier: every rank is a thread on the emulated
    device, NCCL is the thread-backed stand-in of tests/emulator/nccl_threads.cpp (found through SZS_B200_NCCL_LIBRARY, the
    way a real NCCL is). Covers the packed broadcast, the row partition, the grouped gather, symmetric calls with
    asymmetric tables, more ranks than rows, and rejections travelling in the header."""
    from tests.emulator import build as emulator_build

    support.build_oracle()
    environment = dict(os.environ, SZS_B200_LIBRARY=str(emulated_library), SZS_B200_NCCL_LIBRARY=str(emulator_build.build_nccl()),
                       SZB_EMULATED_SMS="2")
    ran = subprocess.run([sys.executable, "tests/emulator/ranks_as_threads.py"], cwd=support.ROOT, env=environment,
                         capture_output=True, text=True, timeout=900)
    assert ran.returncode == 0 and "world 2: ok" in ran.stdout and "world 3: ok" in ran.stdout, (ran.stdout + ran.stderr)[-3000:]
