#!/usr/bin/env python3
"""How much of a workload's roofline gap is STRUCTURE rather than pipeline? (no GPU needed)

Asks the library's own host scheduler (`szs_b200_debug_schedule`, pure host code) for the shapes it gives every candidate
of a bench workload on a 148-SM device, and counts

    useful cells / cells the wavefronts execute

where an executed wavefront of a slot sweeps `K x G` columns (the strip shape; both 16-bit halves of the packed tier) for
`len(query) + G - 1` steps per pass — padding up to the shape's coverage, the fill and drain of the skewed wavefront, the
shorter partner of a packed pair and multi-pass rounding are all in there; the bit-parallel tier sweeps `32 W G` pattern
bits for `len(query) + G - 1` steps. roofline fraction = structural efficiency x what the pipeline achieves on the cells
it executes.

    python tools/structural_efficiency.py c2 c2_weird c3 c3_sw c4 c5
"""
import ctypes as ct
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from stringzilla_b200 import _abi  # noqa: E402

MAX_COVERAGE = 40 * 32


def efficiency(name: str, device_sms: int = 148):
    label, family, kwargs, queries, candidates = bench.make_workload(name)
    engine = bench.make_engine(family, kwargs)
    lib = _abi.bind()
    q = np.array([len(s) for s in queries], dtype=np.uint32)
    c = np.array([len(s) for s in candidates], dtype=np.uint32)
    distinct = len(set(b"".join(candidates[:50])))
    assignment, partner, order = np.zeros(len(c), dtype=np.uint32), np.zeros(len(c), dtype=np.uint32), np.zeros(len(q), dtype=np.uint32)
    n_jobs = ct.c_uint64()
    status = lib.szs_b200_debug_schedule(engine._handle, q.ctypes.data, len(q), c.ctypes.data, len(c), distinct, device_sms,
                                         assignment.ctypes.data, partner.ctypes.data, order.ctypes.data, None, 0, ct.byref(n_jobs))
    assert status == 0
    tier, width, lanes = assignment & 0xFF, (assignment >> 8) & 0xFF, (assignment >> 16) & 0xFF
    useful = float(q.astype(np.float64).sum() * c.astype(np.float64).sum())
    executed = 0.0
    seen = set()
    for cand in range(len(c)):
        if cand in seen:
            continue
        g, w = int(lanes[cand]), int(width[cand])
        steps = (q.astype(np.float64) + g - 1) * (q > 0)
        if tier[cand] == 0:  # bit-parallel: 32 * W * G pattern bits per query and step
            executed += 32.0 * w * g * steps.sum()
            continue
        pair = [cand] + ([int(partner[cand])] if partner[cand] != 0xFFFFFFFF else [])
        seen.update(pair)
        coverage = w * g
        passes = max(1, -(-int(c[pair].max()) // coverage))
        halves = 2 if tier[cand] == 1 else 1  # packed 2 x u16: both halves travel, whether or not a partner fills the second
        executed += float(halves * coverage * passes) * steps.sum()
    return label, useful / executed


if __name__ == "__main__":
    for workload in sys.argv[1:] or ["c2", "c2_weird", "c3", "c3_sw", "c4", "c5"]:
        label, value = efficiency(workload)
        print(f"{value:.3f}  {label}")
