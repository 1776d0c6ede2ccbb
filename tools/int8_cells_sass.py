#!/usr/bin/env python3
"""Why there is no 8-bit cell tier: compiles the inner step of a linear-gap DP cell update —
`max(max(diag + e, up), left)` — for sm_100a once on 4 x u8 lanes (the "video" SIMD intrinsics `__vaddus4` / `__vmaxu4`)
and once on 2 x u16 lanes (DPX, `__vimax3_u16x2`), and prints the SASS instruction count per DP cell. No GPU needed.

    python tools/int8_cells_sass.py
"""
import re
import subprocess
import tempfile
from pathlib import Path

SOURCE = r"""
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void cells_u8x4(uint32_t *out, uint32_t const *in) {
    uint32_t diag = in[threadIdx.x], e = in[threadIdx.x + 32], up = in[threadIdx.x + 64], left = in[threadIdx.x + 96];
    out[threadIdx.x] = __vmaxu4(__vmaxu4(__vaddus4(diag, e), up), left);
}
__global__ void cells_u16x2(uint32_t *out, uint32_t const *in) {
    uint32_t diag = in[threadIdx.x], e = in[threadIdx.x + 32], up = in[threadIdx.x + 64], left = in[threadIdx.x + 96];
    out[threadIdx.x] = __vimax3_u16x2(diag + e, up, left);
}
"""
PLUMBING = ("LDG", "STG", "LDC", "S2R", "S2UR", "IMAD.WIDE", "EXIT", "BRA", "NOP", "MOV", "ULDC", "LDCU")


def main():
    with tempfile.TemporaryDirectory() as tmp:
        source, cubin = Path(tmp) / "cells.cu", Path(tmp) / "cells.cubin"
        source.write_text(SOURCE)
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-cubin", str(source), "-o", str(cubin)],
                       check=True)
        sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], check=True, capture_output=True, text=True).stdout
    for function in sass.split("Function : ")[1:]:
        name = function.split("\n", 1)[0].strip()
        lanes = 4 if "u8x4" in name else 2
        ops = []
        for line in function.splitlines():
            match = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
            if match and not match.group(1).startswith(PLUMBING):
                ops.append(match.group(1))
        print(f"{name}: {len(ops)} arithmetic instructions for {lanes} cells = {len(ops) / lanes:.2f} per cell")
        print("    " + " ".join(ops))


if __name__ == "__main__":
    main()
