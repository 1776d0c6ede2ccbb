#!/usr/bin/env bash
# TEST / MEASUREMENT INFRASTRUCTURE ONLY. Builds the UNMODIFIED reference CUDA library (its CUDA + Kepler + Hopper
# tiers) for sm_100 from the sources under /root/reference into oracle/_ref/libstringzillas_ref_cuda.so, so that
# "the reference kernel on the same box" can be timed next to ours.
# Mirrors the reference's define_stringzillas_cuda_library (CMakeLists.txt:578-690) with one change: the arch list is
# sm_100 (the reference ships only sm_80/sm_90 SASS; its runtime grants the Hopper tier to any sm_code >= 90,
# c/stringzillas/runtime.cuh:41-42). Fingerprint TUs are skipped (out of scope). No reference source is copied.
set -euo pipefail
REF=${SZ_REFERENCE_DIR:-/root/reference}
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
mkdir -p "$OUT/obj_cuda"
if [ ! -d "$REF/c/stringzillas" ]; then
  echo "reference tree not found at $REF; keeping prebuilt $OUT/libstringzillas_ref_cuda.so (if any)"; exit 0
fi
DEFS="-DSZ_DYNAMIC_DISPATCH=1 -DSZ_AVOID_LIBC=0 -DSZ_DEBUG=0 -DSZ_USE_CUDA=1 -DSZ_USE_ROCM=0 -DSZ_IS_64BIT_X86_=1 -DSZ_IS_64BIT_ARM_=0 \
 -DSZ_USE_WESTMERE=1 -DSZ_USE_GOLDMONT=1 -DSZ_USE_HASWELL=1 -DSZ_USE_SKYLAKE=1 -DSZ_USE_ICELAKE=1"
INCS="-I$REF/include -I$REF/forkunion/include"
NVCCFLAGS="-std=c++20 -O2 --expt-relaxed-constexpr -arch=sm_100 -Xcompiler -fPIC,-march=ivybridge -w $DEFS $INCS"
CXXFLAGS="-std=c++20 -O2 -fPIC -march=ivybridge -w $DEFS $INCS"
CU="runtime levenshtein needleman_wunsch smith_waterman levenshtein_cuda needleman_wunsch_cuda smith_waterman_cuda \
 levenshtein_kepler levenshtein_hopper needleman_wunsch_hopper smith_waterman_hopper"
CPP="levenshtein_serial levenshtein_haswell levenshtein_icelake needleman_wunsch_serial needleman_wunsch_haswell \
 needleman_wunsch_icelake smith_waterman_serial smith_waterman_haswell smith_waterman_icelake"
JOBS=${JOBS:-4}
run() { while [ "$(jobs -r | wc -l)" -ge "$JOBS" ]; do sleep 1; done; "$@" & }
for s in $CU; do
  [ "$OUT/obj_cuda/$s.cu.o" -nt "$REF/c/stringzillas/$s.cu" ] || run nvcc $NVCCFLAGS -c "$REF/c/stringzillas/$s.cu" -o "$OUT/obj_cuda/$s.cu.o"
done
for s in $CPP; do
  [ "$OUT/obj_cuda/$s.o" -nt "$REF/c/stringzillas/$s.cpp" ] || run g++ $CXXFLAGS -c "$REF/c/stringzillas/$s.cpp" -o "$OUT/obj_cuda/$s.o"
done
[ "$OUT/obj_cuda/forkunion.o" -nt "$REF/forkunion/c/forkunion.cpp" ] || run g++ $CXXFLAGS -c "$REF/forkunion/c/forkunion.cpp" -o "$OUT/obj_cuda/forkunion.o"
wait
nvcc -arch=sm_100 -shared -o "$OUT/libstringzillas_ref_cuda.so" "$OUT"/obj_cuda/*.o -lcuda -lpthread
echo "built $OUT/libstringzillas_ref_cuda.so"
nm -D --defined-only "$OUT/libstringzillas_ref_cuda.so" | grep -c ' T szs_' || true
