#!/bin/bash
# Memory and undefined-behaviour check of every kernel and of the host driver on the CPU: the GPU suite's parity tests run
# against the emulated library built with AddressSanitizer and UndefinedBehaviorSanitizer (tests/emulator/build.py --sanitize). Device allocations are exact-size heap blocks
# and the part of shared memory a launch did not request is poisoned, so an out-of-bounds access in global OR shared
# memory aborts with a report that names the kernel's source line. Usage: tools/emulator_memcheck.sh [pytest -k expression]
set -e
cd "$(dirname "$0")/.."
LIBRARY=$(python tests/emulator/build.py --sanitize | tail -n 1)
EXPRESSION=${1:-"golden or known_answer or random_cross_products or mixed_lengths or shapes_and_degenerate or symmetric or asymmetric or alphabets or unusual_cost or utf8 or every_input_layout or outside_host_memory or memory_kinds or too_large or randomized_differential or arrow or row_shards or three_shards or bit_parallel_levenshtein_shapes or every_strip or every_word or malformed or alphabet_detection"}
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" UBSAN_OPTIONS=print_stacktrace=1 ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:abort_on_error=1 \
SZS_B200_LIBRARY=$LIBRARY SZB_EMULATED_SMS=2 \
python -m pytest tests/test_parity_gpu.py tests/test_multi_gpu.py -m gpu -q -x -p no:cacheprovider -k "$EXPRESSION"
