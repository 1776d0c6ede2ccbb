#!/usr/bin/env python3
"""Times the UNMODIFIED reference CUDA library rebuilt for sm_100 (oracle/build_ref_cuda.sh ->
oracle/_ref/libstringzillas_ref_cuda.so) on the bench workloads, through the same C ABI and ctypes stub as ours, and
checks that our matrices equal the reference GPU's. Measurement infrastructure only; not part of the product path.

    python tools/ref_cuda_probe.py [workload ...]      # default: c2 c2_weird c3_linear c3 c3_sw
"""
import ctypes as ct
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import stringzilla_b200 as szb  # noqa: E402
from stringzilla_b200 import _abi  # noqa: E402

REF = ROOT / "oracle" / "_ref" / "libstringzillas_ref_cuda.so"


class ManagedTape:
    """A u64 tape whose blob and offsets live in `szs_unified_alloc` memory, as the reference's GPU engines require
    (include/stringzillas/similarities/cuda.cuh:4268-4272)."""

    def __init__(self, lib, tape):
        self.lib = lib
        self.n_bytes, self.n_offsets = max(int(tape.data.size), 1), int(tape.offsets.size)
        self.data_ptr = lib.szs_unified_alloc(self.n_bytes)
        self.offsets_ptr = lib.szs_unified_alloc(self.n_offsets * 8)
        ct.memmove(self.data_ptr, tape.data.ctypes.data, int(tape.data.size))
        ct.memmove(self.offsets_ptr, tape.offsets.ctypes.data, self.n_offsets * 8)
        self.struct = _abi.SequenceU64Tape(self.data_ptr, self.offsets_ptr, len(tape))

    def pointer(self):
        return ct.byref(self.struct)

    def __del__(self):
        self.lib.szs_unified_free(self.data_ptr, self.n_bytes)
        self.lib.szs_unified_free(self.offsets_ptr, self.n_offsets * 8)


def main():
    if not REF.exists():
        print(json.dumps({"impl": "reference-cuda", "unavailable": f"{REF} is not built"}))
        return
    lib = _abi.bind(REF)
    caps = int(lib.szs_capabilities()) & _abi.CAPS_CUDA  # GPU tiers only: the init ladder prefers CPU SIMD tiers
    scope = szb.DeviceScope(gpu_device=0, library_path=REF)
    ours_scope = szb.DeviceScope(gpu_device=0)
    for name in sys.argv[1:] or ["c2", "c2_weird", "c3_linear", "c3", "c3_sw"]:
        label, family, kwargs, queries, candidates = bench.make_workload(name)
        q, c = szb.Tape(queries), szb.Tape(candidates)
        cells = int(q.lengths.sum()) * int(c.lengths.sum())
        engine = bench.make_engine(family, kwargs, library_path=REF, capabilities=caps)
        mq, mc = ManagedTape(lib, q), ManagedTape(lib, c)
        out = np.zeros((len(q), len(c)), dtype=engine._result_dtype)
        function = getattr(lib, f"szs_{engine._family}_u64tape")
        message = ct.c_char_p()
        times = []
        for _ in range(6):
            start = time.perf_counter()
            status = function(engine._handle, scope.handle, mq.pointer(), mc.pointer(), out.ctypes.data, len(c), ct.byref(message))
            times.append(time.perf_counter() - start)
            _abi.check(status, message)
        ours = bench.make_engine(family, kwargs)(q, c, device=ours_scope)
        best = min(times[1:])
        print(json.dumps({"impl": "reference-cuda (sm_100 rebuild)", "workload": label, "gcups_wall_best_of_5": round(cells / best / 1e9, 1),
                          "ms": round(best * 1e3, 3), "capabilities": hex(caps), "matches_ours": bool(np.array_equal(out, ours))}))


if __name__ == "__main__":
    main()
