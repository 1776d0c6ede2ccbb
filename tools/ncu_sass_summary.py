#!/usr/bin/env python3
"""SASS-level summary of the hottest loop of every kernel in an .ncu-rep captured with `--set full --import-source on`:
how many instructions a step has, how the warp samples split over stall reasons, and which opcodes collect them.
Usage: python tools/ncu_sass_summary.py gpurun_out/prof.ncu-rep > profiles/<name>_sass.txt"""
import collections
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kernels, current = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        current = {"name": row[1], "rows": []}
        kernels.append(current)
    elif current is not None:
        current["rows"].append(row)
print(f"# {sys.argv[1]}: SASS-level view of each kernel's hot loop (`ncu -i ... --page source --csv --print-source sass`)\n")
seen = set()
for kernel in kernels:
    header = kernel["rows"][0]
    index = {name: i for i, name in enumerate(header)}
    data = [r for r in kernel["rows"][1:] if len(r) == len(header)]
    if not data or kernel["name"] in seen:
        continue
    seen.add(kernel["name"])
    most = max(int(r[index["Instructions Executed"]]) for r in data)
    hot = [r for r in data if int(r[index["Instructions Executed"]]) > 0.45 * most]
    total = sum(int(r[index["# Samples"]]) for r in data)
    hot_samples = sum(int(r[index["# Samples"]]) for r in hot)
    print(f"## {kernel['name']}")
    print(f"hot loop: {len(hot)} SASS instructions, each executed {most} times; {hot_samples / max(total, 1):.1%} of all warp samples")
    states = collections.Counter()
    for r in hot:
        for name in header:
            if name.startswith("stall_") and "Not Issued" not in name:
                states[name] += int(r[index[name]])
    everything = sum(states.values()) or 1
    print("warp samples in the hot loop by state: " + ", ".join(f"{s[6:]} {v / everything:.1%}" for s, v in states.most_common(7)))
    by_opcode, count = collections.Counter(), collections.Counter()
    for r in hot:
        words = r[index["Source"]].split()
        opcode = (words[1] if words[0].startswith("@") else words[0]).split(".")[0]
        by_opcode[opcode] += int(r[index["# Samples"]])
        count[opcode] += 1
    print("instructions per step / share of samples: " +
          ", ".join(f"{o} {count[o]} / {by_opcode[o] / max(hot_samples, 1):.1%}" for o, _ in by_opcode.most_common(8)))
    print()
