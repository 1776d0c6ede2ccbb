#!/usr/bin/env bash
# TEST / MEASUREMENT INFRASTRUCTURE ONLY. Builds the REFERENCE's own CUPS benchmark
# (bench/similarities.cu) with this repo's engines registered inside it: a patched copy of bench/similarities.cuh is
# written to oracle/_ref/bench_src/ (oracle/patch_ref_bench.py), compiled against the reference's headers and linked
# against stringzilla_b200/lib/libstringzillas_b200.so. The reference's serial baselines, cell accounting, equality check
# (bench/similarities.cuh:410-423) and "Kernel:" line then cover `*_b200_*` engines next to its own `*_cuda_*` ones.
# Output: oracle/_ref/ref_bench_similarities. No reference source enters the repository (oracle/_ref is git-ignored).
set -euo pipefail
REF=${SZ_REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
OUT="$HERE/_ref"
if [ ! -d "$REF/bench" ]; then
  echo "reference tree not found at $REF; keeping prebuilt $OUT/ref_bench_similarities (if any)"; exit 0
fi
mkdir -p "$OUT/bench_src"
cp "$REF/bench/similarities.cu" "$REF/bench/shared.hpp" "$OUT/bench_src/"
python3 "$HERE/patch_ref_bench.py" "$REF/bench/similarities.cuh" "$OUT/bench_src/similarities.cuh"
DEFS="-DSZ_DYNAMIC_DISPATCH=1 -DSZ_AVOID_LIBC=0 -DSZ_DEBUG=0 -DSZ_USE_CUDA=1 -DSZ_USE_ROCM=0 -DSZ_IS_64BIT_X86_=1 -DSZ_IS_64BIT_ARM_=0 \
 -DSZ_USE_WESTMERE=1 -DSZ_USE_GOLDMONT=1 -DSZ_USE_HASWELL=1 -DSZ_USE_SKYLAKE=1 -DSZ_USE_ICELAKE=1"
INCS="-I$OUT/bench_src -I$ROOT/bench -I$REF/include -I$REF/forkunion/include -I$REF/test"
if [ ! "$OUT/bench_src/similarities.o" -nt "$ROOT/bench/reference_bench_b200.cuh" ]; then
  nvcc -std=c++20 -O2 --expt-relaxed-constexpr -arch=sm_100 -Xcompiler -march=ivybridge -w $DEFS $INCS \
    -c "$OUT/bench_src/similarities.cu" -o "$OUT/bench_src/similarities.o"
fi
# The reference's engine templates are `extern template`s in the benchmark: their explicit instantiations come from
# the reference's per-ISA units (built by oracle/build_ref_cuda.sh into oracle/_ref/obj_cuda). The reference's four
# C-ABI units (runtime, levenshtein, needleman_wunsch, smith_waterman .cu) are deliberately NOT linked: every `szs_*`
# symbol the adapters call must resolve to this repo's library.
[ -f "$OUT/obj_cuda/levenshtein_cuda.cu.o" ] || bash "$HERE/build_ref_cuda.sh"
UNITS=""
for s in levenshtein_cuda needleman_wunsch_cuda smith_waterman_cuda levenshtein_kepler levenshtein_hopper \
         needleman_wunsch_hopper smith_waterman_hopper; do UNITS="$UNITS $OUT/obj_cuda/$s.cu.o"; done
for s in levenshtein_serial levenshtein_haswell levenshtein_icelake needleman_wunsch_serial needleman_wunsch_haswell \
         needleman_wunsch_icelake smith_waterman_serial smith_waterman_haswell smith_waterman_icelake forkunion; do
  UNITS="$UNITS $OUT/obj_cuda/$s.o"; done
nvcc -arch=sm_100 -o "$OUT/ref_bench_similarities" "$OUT/bench_src/similarities.o" $UNITS \
  -L"$ROOT/stringzilla_b200/lib" -lstringzillas_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../stringzilla_b200/lib' -lcuda -lpthread
echo "built $OUT/ref_bench_similarities"
