#!/usr/bin/env bash
# TEST / MEASUREMENT INFRASTRUCTURE ONLY. Builds the REFERENCE's two CPython
# extensions from the sources under /root/reference, unmodified:
#   oracle/_ref/python/stringzilla<ext>   the single-string module (provides the `Strs` capsule the binding imports)
#   oracle/_ref/python/stringzillas<ext>  python/stringzillas/*.c — the reference's binding of the similarity engines —
#                                          LINKED AGAINST stringzilla_b200/lib/libstringzillas_b200.so instead of the
#                                          reference's own C shims (no c/stringzillas/*.cu is compiled in).
# tests/test_reference_python_binding.py then runs the reference binding's API on the GPU. Nothing is copied into the repo.
set -euo pipefail
REF=${SZ_REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
OUT="$HERE/_ref/python"
if [ ! -d "$REF/python/stringzillas" ]; then
  echo "reference tree not found at $REF; keeping prebuilt $OUT (if any)"; exit 0
fi
mkdir -p "$OUT/obj"
PY=${PYTHON:-python3}
PYINC=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
NPINC=$($PY -c "import numpy; print(numpy.get_include())")
SUFFIX=$($PY -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
MACROS="-DSZ_DYNAMIC_DISPATCH=1 -DSZ_IS_BIG_ENDIAN_=0 -DSZ_IS_64BIT_X86_=1 -DSZ_IS_64BIT_ARM_=0 -DSZ_USE_WESTMERE=1 -DSZ_USE_GOLDMONT=1 \
 -DSZ_USE_HASWELL=1 -DSZ_USE_SKYLAKE=1 -DSZ_USE_ICELAKE=1 -DSZ_USE_NEON=0 -DSZ_USE_NEONAES=0 -DSZ_USE_NEONSHA=0 -DSZ_USE_SVE=0 \
 -DSZ_USE_SVE2=0 -DSZ_USE_SVE2AES=0"
JOBS=${JOBS:-8}
run() { while [ "$(jobs -r | wc -l)" -ge "$JOBS" ]; do sleep 0.2; done; "$@" & }

# 1. The single-string module (mirrors setup.py:733-770: C99, -O2, every x86 tier stamped on, run-time dispatch).
if [ ! -f "$OUT/stringzilla$SUFFIX" ]; then
  for f in "$REF"/python/stringzilla/*.c; do
    run gcc -std=c99 -D_GNU_SOURCE -O2 -fPIC -w $MACROS -I"$REF/include" -I"$REF/c/stringzilla" -I"$PYINC" -c "$f" -o "$OUT/obj/py_$(basename "$f" .c).o"
  done
  for f in "$REF"/c/stringzilla/*.c; do
    run gcc -std=c99 -D_GNU_SOURCE -O2 -fPIC -w $MACROS -I"$REF/include" -I"$REF/c/stringzilla" -c "$f" -o "$OUT/obj/c_$(basename "$f" .c).o"
  done
  wait
  gcc -shared -fPIC -o "$OUT/stringzilla$SUFFIX" "$OUT"/obj/py_*.o "$OUT"/obj/c_*.o -lm
fi

# 2. The similarity binding, against THIS repo's library (mirrors setup.py:772-790 minus every c/stringzillas source).
for f in stringzillas device_scope similarities fingerprints; do
  run gcc -std=c99 -D_GNU_SOURCE -O2 -fPIC -w $MACROS -DSZ_USE_CUDA=1 -I"$REF/include" -I"$REF/python" -I"$REF/c/stringzillas" \
    -I"$REF/forkunion/include" -I"$PYINC" -I"$NPINC" -c "$REF/python/stringzillas/$f.c" -o "$OUT/obj/szs_$f.o"
done
wait
g++ -shared -fPIC -o "$OUT/stringzillas$SUFFIX" "$OUT"/obj/szs_*.o -L"$ROOT/stringzilla_b200/lib" -lstringzillas_b200 \
  -Wl,-rpath,'$ORIGIN/../../../stringzilla_b200/lib'
echo "built $OUT/stringzilla$SUFFIX and $OUT/stringzillas$SUFFIX"
