#!/usr/bin/env python3
"""TEST / MEASUREMENT INFRASTRUCTURE ONLY. Writes a patched COPY of the reference's benchmark header
(/root/reference/bench/similarities.cuh) into oracle/_ref/bench_src/ (git-ignored): every `#if SZ_USE_CUDA` block
that registers engines with `bench_unary` is followed by a copy of itself in which the reference's CUDA engine types
are replaced by the adapters of bench/reference_bench_b200.cuh (`levenshtein_cuda_t` -> `b200_adapters::levenshtein_b200_t`,
label `levenshtein_cuda_` -> `levenshtein_b200_`, ...). Nothing of the reference is copied into this repository: the
copy lives under oracle/_ref/ only, and the registrations are derived mechanically from the reference's own.

    python oracle/patch_ref_bench.py <reference similarities.cuh> <output similarities.cuh>
"""
import re
import sys

source, target = sys.argv[1], sys.argv[2]
lines = open(source).read().split("\n")
out, i, spliced = [], 0, 0
while i < len(lines):
    line = lines[i]
    out.append(line)
    if line.strip() == '#include "shared.hpp"':
        out.append('#include "reference_bench_b200.cuh" // adapters over libstringzillas_b200.so (this repo)')
    if line.strip() == "#if SZ_USE_CUDA":
        depth, j = 1, i + 1
        while j < len(lines) and depth:
            stripped = lines[j].strip()
            if stripped.startswith("#if"):
                depth += 1
            elif stripped.startswith("#endif"):
                depth -= 1
            j += 1
        block = lines[i:j]
        out.extend(lines[i + 1 : j])
        i = j
        if any("bench_unary(" in b for b in block):
            copy = "\n".join(block)
            copy = re.sub(r"\b(\w+)_cuda_t\b", r"b200_adapters::\1_b200_t", copy)
            copy = re.sub(r'_cuda(_"s|:"s)', r"_b200\1", copy)
            out.append("// --- spliced by oracle/patch_ref_bench.py: the same registrations over libstringzillas_b200.so ---")
            out.extend(copy.split("\n"))
            spliced += 1
        continue
    i += 1
open(target, "w").write("\n".join(out))
print(f"spliced {spliced} registration block(s) into {target}")
if spliced < 2:
    sys.exit("expected the Levenshtein and the Needleman-Wunsch / Smith-Waterman registration blocks")
