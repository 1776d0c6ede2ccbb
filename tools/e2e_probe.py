"""Scratch probe: where does a host-buffer call spend its wall time?"""
import ctypes as ct, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import stringzilla_b200 as szb
from stringzilla_b200 import _abi
import bench

label, family, kwargs, queries, candidates = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "c2")
q, c = szb.Tape(queries), szb.Tape(candidates)
engine = bench.make_engine(family, kwargs)
scope = szb.DeviceScope(gpu_device=0)
out = np.zeros((len(q), len(c)), dtype=engine._result_dtype)
for kind in ("pageable", "pinned", "device"):
    if kind == "pinned":
        qd, cd = torch.from_numpy(q.data).pin_memory(), torch.from_numpy(c.data).pin_memory()
        od = torch.zeros((len(q), len(c)), dtype=torch.int64).pin_memory()
    elif kind == "device":
        qd, cd = torch.from_numpy(q.data).cuda(), torch.from_numpy(c.data).cuda()
        od = torch.zeros((len(q), len(c)), dtype=torch.int64, device="cuda")
    else:
        qd, cd, od = torch.from_numpy(q.data), torch.from_numpy(c.data), torch.zeros((len(q), len(c)), dtype=torch.int64)
    tq = _abi.SequenceU64Tape(qd.data_ptr(), q.offsets.ctypes.data, len(q))
    tc = _abi.SequenceU64Tape(cd.data_ptr(), c.offsets.ctypes.data, len(c))
    msg = ct.c_char_p()
    fn = getattr(_abi.bind(), f"szs_{engine._family}_u64tape")
    for i in range(6):
        t0 = time.perf_counter()
        st = fn(engine._handle, scope.handle, ct.byref(tq), ct.byref(tc), od.data_ptr(), len(c), ct.byref(msg))
        wall = (time.perf_counter() - t0) * 1e3
        s = engine.last_call_stats
        print(f"{kind:9s} call {i}: wall {wall:8.3f} ms  device-total {s.total_milliseconds:8.3f}  kernels {s.kernel_milliseconds:8.3f}  launches {s.kernel_launches}  status {st}")
