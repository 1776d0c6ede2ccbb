#!/bin/bash
# One GPU call: A/B of the scheduling experiments on c2, then the GPU parity suite and the bench line through the winner.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/shot_gpu.txt 2>&1
timeout 200 python tools/ab_variants.py c2 base=stringzilla_b200/lib share=stringzilla_b200/lib_share \
    share_g8=stringzilla_b200/lib_share_g8 g8=stringzilla_b200/lib_g8 --reps 20 > gpurun_out/ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/ab.log
W=$(python - <<'P'
import json
best = None
try:
    for line in open('gpurun_out/ab_variants.jsonl'):
        r = json.loads(line)
        if 'fastest_parity_clean_build' in r:
            best = r['fastest_parity_clean_build']
except OSError:
    pass
print(best or 'base')
P
)
DIR=stringzilla_b200/lib; [ "$W" != base ] && DIR=stringzilla_b200/lib_$W
echo "winner $W -> $DIR" >> gpurun_out/ab.log
export SZS_B200_LIBRARY=$PWD/$DIR/libstringzillas_b200.so
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_winner.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_winner.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_winner.json 2> gpurun_out/bench_winner.err
echo "bench rc=$?" >> gpurun_out/bench_winner.err
cat gpurun_out/ab.log; tail -4 gpurun_out/pytest_winner.log; tail -2 gpurun_out/bench_winner.err
