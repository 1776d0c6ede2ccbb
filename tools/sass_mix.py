#!/usr/bin/env python3
"""Kernel-tuning aid (no GPU needed): compiles ONE wavefront kernel instantiation (family, strip index) or ONE
bit-parallel kernel (words) to SASS and prints the instruction mix of its hottest loop — the loop body with the most
DPX / LOP3 instructions — so that instruction-count experiments can be judged before spending GPU time.

    python tools/sass_mix.py wavefront <family 0..8> <strip index 0..8> [extra nvcc flags...]
    python tools/sass_mix.py myers <words 1..10> [extra nvcc flags...]
"""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "stringzilla_b200" / "csrc"
TOP = 2
ALU = ("LOP3", "IADD3", "SHF", "VIADDMNMX", "VIMNMX3", "VIMNMX", "ISETP", "SEL", "PRMT", "VIADD", "PLOP3", "IABS", "LEA", "POPC", "FLO", "VABSDIFF")
FMA = ("IMAD", "FFMA", "FMUL", "FADD")


def disassemble(kind, a, b, extra):
    with tempfile.TemporaryDirectory() as tmp:
        out = Path(tmp) / "k.cubin"
        if kind == "wavefront":
            cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-cubin", f"-DSZB_FAMILY={a}",
                   f"-DSZB_ONLY_STRIP={b}", "-diag-suppress", "177", *extra, str(CSRC / "family.cu"), "-o", str(out)]
        else:
            cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-cubin", f"-DSZB_ONLY_WORDS={a}",
                   "-diag-suppress", "177", *extra, str(CSRC / "myers.cu"), "-o", str(out)]
        subprocess.run(cmd, check=True)
        return subprocess.run(["cuobjdump", "-sass", str(out)], check=True, capture_output=True, text=True).stdout


def main():
    kind = sys.argv[1]
    if kind == "wavefront":
        a, b, extra = sys.argv[2], sys.argv[3], sys.argv[4:]
    else:
        a, b, extra = sys.argv[2], None, sys.argv[3:]
    text = disassemble(kind, a, b, extra)
    for function in text.split("Function : ")[1:]:
        name = function.split("\n", 1)[0]
        instructions = []
        for line in function.splitlines():
            match = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if match:
                instructions.append((int(match.group(1), 16), match.group(2).strip()))
        address_index = {address: i for i, (address, _) in enumerate(instructions)}
        # Loops = backward branches; the hottest is the one whose body holds the most DPX / LOP3 work per instruction span.
        loops = []
        for i, (address, text_) in enumerate(instructions):
            match = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", text_)
            if not match:
                continue
            target = int(match.group(1), 16)
            if target >= address or target not in address_index:
                continue
            loops.append((address_index[target], i))
        ranked = []
        for first, last in loops:
            if any((f, l) != (first, last) and f >= first and l <= last for f, l in loops):
                continue  # encloses another loop
            body = instructions[first : last + 1]
            ranked.append((sum(1 for _, t in body if re.search(r"VIADDMNMX|VIMNMX|LOP3", t)), body))
        ranked.sort(key=lambda item: -item[0])
        for hot, body in ranked[:TOP]:
            report(name, body)


def report(name, body):
    if True:
        mix = collections.Counter()
        for _, t in body:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            mix[t.split()[0].split(".")[0]] += 1
        alu = sum(n for op, n in mix.items() if op in ALU)
        fma = sum(n for op, n in mix.items() if op in FMA)
        movs = sum(1 for _, t in body if "IMAD.MOV" in t or re.match(r"(@!?P\d\s+)?MOV\b", t))
        print(f"{name.strip()}\n  hottest loop: {len(body)} instructions, ALU-pipe {alu}, FMA-pipe {fma} (of which moves {movs}), "
              f"LDS {mix.get('LDS', 0)}, SHFL {mix.get('SHFL', 0)}")
        print("  " + ", ".join(f"{op} {n}" for op, n in mix.most_common()))


if __name__ == "__main__":
    main()
