#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY. Builds the UNMODIFIED reference CPU library (StringZilla 5.1.2 `stringzillas` layer)
# from the sources where they lie under /root/reference into oracle/_ref/libstringzillas_ref.so.
#
# Recipe mirrors the reference's own `define_stringzillas_shared` target (CMakeLists.txt:512-552, sources :605-630):
#   - common definitions SZ_DYNAMIC_DISPATCH=1 SZ_AVOID_LIBC=0 SZ_DEBUG=0 (CMakeLists.txt:506), SZ_USE_CUDA=0
#   - -O2 -fPIC (CMakeLists.txt:548) and the conservative `ivybridge` baseline (cmake/sz_compiler_flags.cmake:318-324):
#     the AVX2/AVX-512 kernels carry per-function target attributes and are selected at run time from the
#     capability mask, so the library is portable to whatever x86 host the GPU box has.
#   - every x86 capability the toolchain can compile is stamped on (set_architecture_simd_definitions COMPILE).
# No reference source is copied into this repo; outputs go only to oracle/_ref/ (git-ignored, travels via gpurun).
set -euo pipefail
REF=${SZ_REFERENCE_DIR:-/root/reference}
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
mkdir -p "$OUT/obj"
if [ ! -d "$REF/c/stringzillas" ]; then
  echo "reference tree not found at $REF; keeping prebuilt $OUT/libstringzillas_ref.so (if any)"; exit 0
fi
CXX=${CXX:-g++}
FLAGS="-std=c++20 -O2 -fPIC -march=ivybridge -DSZ_DYNAMIC_DISPATCH=1 -DSZ_AVOID_LIBC=0 -DSZ_DEBUG=0 -DSZ_USE_CUDA=0 -DSZ_USE_ROCM=0 \
 -DSZ_IS_64BIT_X86_=1 -DSZ_IS_64BIT_ARM_=0 -DSZ_USE_WESTMERE=1 -DSZ_USE_GOLDMONT=1 -DSZ_USE_HASWELL=1 -DSZ_USE_SKYLAKE=1 -DSZ_USE_ICELAKE=1 \
 -I$REF/include -I$REF/forkunion/include -w"
SRCS="runtime levenshtein needleman_wunsch smith_waterman \
 levenshtein_serial levenshtein_haswell levenshtein_icelake \
 needleman_wunsch_serial needleman_wunsch_haswell needleman_wunsch_icelake \
 smith_waterman_serial smith_waterman_haswell smith_waterman_icelake"
pids=()
for s in $SRCS; do
  if [ "$OUT/obj/$s.o" -nt "$REF/c/stringzillas/$s.cpp" ]; then continue; fi
  $CXX $FLAGS -c "$REF/c/stringzillas/$s.cpp" -o "$OUT/obj/$s.o" &
  pids+=($!)
done
if [ ! "$OUT/obj/forkunion.o" -nt "$REF/forkunion/c/forkunion.cpp" ]; then
  $CXX $FLAGS -c "$REF/forkunion/c/forkunion.cpp" -o "$OUT/obj/forkunion.o" & pids+=($!)
fi
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$CXX -shared -o "$OUT/libstringzillas_ref.so" "$OUT"/obj/*.o -lpthread
echo "built $OUT/libstringzillas_ref.so"
nm -D --defined-only "$OUT/libstringzillas_ref.so" | grep -c ' T szs_' || true
