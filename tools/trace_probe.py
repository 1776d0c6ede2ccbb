import sys, os
sys.path.insert(0, "/root/repo")
import ctypes as ct, numpy as np, torch
import stringzilla_b200 as szb
from stringzilla_b200 import _abi
import bench
for name in ("s100_lev", "c2"):
    label, family, kwargs, queries, candidates = bench.make_workload(name)
    q, c = szb.Tape(queries), szb.Tape(candidates)
    engine = bench.make_engine(family, kwargs); scope = szb.DeviceScope(gpu_device=0)
    qd, cd = torch.from_numpy(q.data).cuda(), torch.from_numpy(c.data).cuda()
    od = torch.zeros((len(q), len(c)), dtype=torch.int64, device="cuda")
    tq = _abi.SequenceU64Tape(qd.data_ptr(), q.offsets.ctypes.data, len(q)); tc = _abi.SequenceU64Tape(cd.data_ptr(), c.offsets.ctypes.data, len(c))
    msg = ct.c_char_p(); fn = getattr(_abi.bind(), f"szs_{engine._family}_u64tape")
    for i in range(3):
        if i == 2: os.environ["SZB200_TRACE"] = "1"; print("==", name, file=sys.stderr)
        fn(engine._handle, scope.handle, ct.byref(tq), ct.byref(tc), od.data_ptr(), len(c), ct.byref(msg))
    os.environ.pop("SZB200_TRACE", None)
    s = engine.last_call_stats; print(name, "device total", s.total_milliseconds, "kernels", s.kernel_milliseconds, file=sys.stderr)
