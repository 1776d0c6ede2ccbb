#!/usr/bin/env python3
"""Writes a synthetic line-per-string dataset for the reference's benchmarks (STRINGWARS_DATASET + STRINGWARS_TOKENS=lines):
    python tools/make_dataset.py <path> <count> <min length> <max length> <dna|protein> [seed]"""
import sys

import numpy as np

path, count, low, high, kind = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
rng = np.random.default_rng(int(sys.argv[6]) if len(sys.argv) > 6 else 42)
alphabet = np.frombuffer(b"ACGT" if kind == "dna" else b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
with open(path, "wb") as out:
    for length in rng.integers(low, high + 1, size=count):
        out.write(alphabet[rng.integers(0, len(alphabet), size=int(length))].tobytes() + b"\n")
