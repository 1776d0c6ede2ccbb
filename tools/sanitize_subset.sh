#!/usr/bin/env bash
# compute-sanitizer over a small slice of the GPU suite (memcheck, then racecheck): the shared-memory profile / match
# tables, the TMA-staged candidate windows (mbarrier), the rune hash table and the multi-pass boundary scratch are exactly
# where a race or an out-of-bounds access would hide. The multi-GPU orchestrator runs its shards on one device.
set -u
export PYTHONUNBUFFERED=1
SELECT='test_known_answer or test_shapes_and_degenerate or test_symmetric_self or test_utf8 or (test_randomized_differential and (0 or 1)) or test_bit_parallel_levenshtein_large or test_mixed_lengths or test_retrieval_shapes'
# racecheck slows kernels down by two to three orders of magnitude: it gets the small-input tests only, and a deadline.
RACE_SELECT='test_known_answer or test_shapes_and_degenerate or test_symmetric_self or test_utf8'
for tool in ${TOOLS:-memcheck racecheck}; do
  echo "== compute-sanitizer --tool $tool"
  if [ $tool = racecheck ]; then pick="$RACE_SELECT"; else pick="$SELECT or test_sharded_call"; fi
  timeout ${SANITIZER_DEADLINE:-420} compute-sanitizer --tool $tool --target-processes all --error-exitcode 99 \
    python -m pytest tests/test_parity_gpu.py tests/test_multi_gpu.py -m gpu -x -q -k "$pick" 2>&1 | tail -15
  echo "exit: ${PIPESTATUS[0]}"
done
