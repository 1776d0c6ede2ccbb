#!/usr/bin/env python3
"""Kernel / scheduling experiments: times ONE bench workload through several builds of the library in one process and
checks every build's matrix against the reference CPU library (oracle/_ref, checker only).

    make -C stringzilla_b200/csrc OUT=../lib_share EXTRA=-DSZB_SHARE_RESIDENT_SLOTS=1
    python tools/ab_variants.py c2 base=stringzilla_b200/lib share=stringzilla_b200/lib_share [--reps 15]

Per build: the library's own CUDA-event kernel time with inputs and results resident in HBM (median and best of the
repetitions), the wall time of the host-buffer call (pinned memory), and `parity`. One JSON line per build, appended to
gpurun_out/ab_variants.jsonl as soon as it is known; the last line names the fastest parity-clean build.
"""
import ctypes as ct
import json
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import stringzilla_b200 as szb  # noqa: E402
from stringzilla_b200 import _abi  # noqa: E402


def main():
    arguments = [a for a in sys.argv[1:] if not a.startswith("--")]
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 15
    workload, builds = arguments[0], [a.split("=", 1) for a in arguments[1:] if "=" in a]
    label, family, kwargs, queries, candidates = bench.make_workload(workload)
    q, c = szb.Tape(queries), szb.Tape(candidates)
    cells = int(q.lengths.sum()) * int(c.lengths.sum())
    rows = bench.PARITY_SAMPLE_ROWS.get(workload)
    sample = np.arange(len(q)) if rows is None else np.linspace(0, len(q) - 1, rows).astype(np.int64)
    want = bench.reference_rows(family, kwargs, queries, candidates, sample)
    log = ROOT / "gpurun_out" / "ab_variants.jsonl"
    log.parent.mkdir(exist_ok=True)
    device = torch.device("cuda", 0)
    q_dev, c_dev = torch.from_numpy(q.data).to(device), torch.from_numpy(c.data).to(device)
    q_pin, c_pin = torch.from_numpy(q.data).pin_memory(), torch.from_numpy(c.data).pin_memory()
    best = None
    for name, directory in builds:
        path = (ROOT / directory / "libstringzillas_b200.so").resolve()
        lib = _abi.bind(path)
        engine = bench.make_engine(family, kwargs, library_path=path)
        scope = szb.DeviceScope(gpu_device=0, library_path=path)
        out_dev = torch.zeros((len(q), len(c)), dtype=torch.int64, device=device)
        out_pin = torch.zeros((len(q), len(c)), dtype=torch.int64).pin_memory()
        message = ct.c_char_p()
        call = getattr(lib, f"szs_{engine._family}_u64tape")

        def run(data_q, data_c, out):
            tq = _abi.SequenceU64Tape(data_q.data_ptr(), q.offsets.ctypes.data, len(q))
            tc = _abi.SequenceU64Tape(data_c.data_ptr(), c.offsets.ctypes.data, len(c))
            _abi.check(call(engine._handle, scope.handle, ct.byref(tq), ct.byref(tc), out.data_ptr(), len(c), ct.byref(message)), message)

        kernel_ms, wall_ms = [], []
        for i in range(3 + reps):
            run(q_dev, c_dev, out_dev)
            if i >= 3:
                kernel_ms.append(engine.last_call_stats.kernel_milliseconds)
        for i in range(3 + reps):
            start = time.perf_counter()
            run(q_pin, c_pin, out_pin)
            if i >= 3:
                wall_ms.append((time.perf_counter() - start) * 1e3)
        got_resident = out_dev.cpu().numpy().view(engine._result_dtype)[sample]
        got_host = out_pin.numpy().view(engine._result_dtype)[sample]
        parity = bool(np.array_equal(got_resident, want) and np.array_equal(got_host, want))
        record = {"workload": label, "build": name, "parity": parity, "launches": engine.last_call_stats.kernel_launches,
                  "kernel_ms_median": round(statistics.median(kernel_ms), 4), "kernel_ms_best": round(min(kernel_ms), 4),
                  "kernel_gcups_median": round(cells / statistics.median(kernel_ms) / 1e6, 1),
                  "e2e_ms_median": round(statistics.median(wall_ms), 4), "e2e_gcups_median": round(cells / statistics.median(wall_ms) / 1e6, 1)}
        print(json.dumps(record), flush=True)
        with open(log, "a") as handle:
            handle.write(json.dumps(record) + "\n")
        # The first build listed is the incumbent: a challenger has to beat it by more than run-to-run noise (1.5 %).
        if parity and (best is None or record["kernel_ms_median"] < 0.985 * best["kernel_ms_median"]):
            best = record
        del engine, scope
    verdict = {"workload": label, "fastest_parity_clean_build": best["build"] if best else None}
    print(json.dumps(verdict), flush=True)
    with open(log, "a") as handle:
        handle.write(json.dumps(verdict) + "\n")


if __name__ == "__main__":
    main()
