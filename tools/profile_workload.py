#!/usr/bin/env python3
"""Runs one bench workload a few times through the library (device-resident inputs) — the target of `ncu` captures:

    ncu --set full --clock-control none --import-source on -k regex:'myers_kernel|wavefront_kernel' -c 4 \
        -o gpurun_out/prof_<name> python tools/profile_workload.py <workload> [calls] [bit_parallel=0]
"""
import ctypes as ct
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
import stringzilla_b200 as szb  # noqa: E402
from stringzilla_b200 import _abi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
label, family, kwargs, queries, candidates = bench.make_workload(name)
if len(sys.argv) > 3 and sys.argv[3] == "0":
    kwargs = dict(kwargs, bit_parallel=False)
engine = bench.make_engine(family, kwargs)
scope = szb.DeviceScope(gpu_device=0)
lib = _abi.bind()
q, c = szb.Tape(queries), szb.Tape(candidates)
device = torch.device("cuda", 0)
q_dev, c_dev = torch.from_numpy(q.data).to(device), torch.from_numpy(c.data).to(device)
out = torch.zeros((len(q), len(c)), dtype=torch.int64, device=device)
tq = _abi.SequenceU64Tape(q_dev.data_ptr(), q.offsets.ctypes.data, len(q))
tc = _abi.SequenceU64Tape(c_dev.data_ptr(), c.offsets.ctypes.data, len(c))
message = ct.c_char_p()
for _ in range(calls):
    _abi.check(lib.szs_b200_scores_u64tape_rows(engine._family_index, engine._handle, scope.handle, ct.byref(tq), ct.byref(tc),
                                                0, len(q), out.data_ptr(), len(c), ct.byref(message)), message)
stats = engine.last_call_stats
print(f"{label}: kernel {stats.kernel_milliseconds:.3f} ms, {stats.cells / stats.kernel_milliseconds / 1e6:.0f} GCUPS, "
      f"{stats.kernel_launches} launches")
