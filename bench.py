#!/usr/bin/env python3
"""bench.py — cell-updates/s of the batched string-similarity hot path, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

A *step* is one pass of the hot path over one synthetic tile: a single `szs_*_u64tape` cross-product call.
The default workload is BASELINE.json `configs[1]`: ~100k (query, candidate) pairs of DNA strings with lengths
uniform on [960, 1088], unit-cost Levenshtein — the tile `q316 x c317` of SURVEY.md §8(d).

Numbers on the line
  value        GCUPS with the strings already resident in HBM and the score matrix left in HBM; time = CUDA events
               recorded by the library on the stream its kernels run on (staging D2D + scheduling + kernels + D2D),
               max over ranks.
  e2e          the same metric through the C-ABI call with pinned HOST buffers: H2D of the strings and D2H of the
               score matrix inside the timed region (wall clock around the synchronous calls, max over ranks).
  roofline     the wavefront kernel against the int16-DPX issue roofline measured live by tools/microbench/pipe_rates.
  cpu_baseline the reference's own CPU library (oracle/_ref) on this box's host cores, same tile.
N > 1 (launched by torch.distributed.run): weak scaling — the tile grows to 316*N query rows, rank r scores its row
block (balanced by total query length); the e2e leg broadcasts the inputs from rank 0 and gathers the score slabs there
over NCCL.
"""

from __future__ import annotations

import argparse
import ctypes as ct
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DNA = b"ACGT"
PROTEIN = b"ARNDCQEGHILKMFPSTWYV"


# ---------------------------------------------------------------------------------------------------------------------
# Workloads (fixed seed 42)
# ---------------------------------------------------------------------------------------------------------------------


def random_strings(rng, count, low, high, alphabet, log_uniform=False):
    symbols = np.frombuffer(alphabet, dtype=np.uint8)
    if log_uniform:
        lengths = np.exp(rng.uniform(np.log(low), np.log(high), size=count)).astype(np.int64)
    else:
        lengths = rng.integers(low, high + 1, size=count)
    return [symbols[rng.integers(0, len(symbols), size=int(n))].tobytes() for n in lengths]


def make_workload(name: str, scale: int = 1):
    """Returns (description, family, engine_kwargs, queries, candidates). `scale` multiplies the query rows (weak scaling)."""
    from stringzilla_b200 import scoring

    rng = np.random.default_rng(42)
    if name in ("c2", "c2_weird", "c1", "c2_text", "c2_text_weird"):
        if name == "c1":
            rows, cols, low, high = 32, 32, 64, 256
        else:
            rows, cols, low, high = 316, 317, 960, 1088
        # "c2_text": the same tile over a 64-symbol alphabet, the regime of natural-language bytes (> 31 classes).
        alphabet = DNA if not name.startswith("c2_text") else bytes(range(48, 112))
        queries = random_strings(rng, rows * scale, low, high, alphabet)
        candidates = random_strings(rng, cols, low, high, alphabet)
        costs = dict(match=0, mismatch=1, open=1, extend=1) if "weird" not in name else dict(match=1, mismatch=3, open=3, extend=3)
        kind = "DNA" if not name.startswith("c2_text") else "64-symbol text"
        label = f"{name}: q{rows * scale} x c{cols} {kind}, len U[{low},{high}], Levenshtein costs {tuple(costs.values())}"
        return label, "levenshtein", costs, queries, candidates
    if name in ("c3", "c3_linear", "c3_sw", "c4", "c5"):
        byte_to_class, table = scoring.blosum62()
        if name == "c4":
            rows, cols = 1000, 1000
            queries = random_strings(rng, rows * scale, 32, 4096, PROTEIN, log_uniform=True)
            candidates = random_strings(rng, cols, 32, 4096, PROTEIN, log_uniform=True)
            family = "smith_waterman"
        elif name == "c5":
            rows, cols = 10000, 10000
            queries = random_strings(rng, rows * scale, 480, 544, PROTEIN)
            candidates = random_strings(rng, cols, 480, 544, PROTEIN)
            family = "needleman_wunsch"
        else:
            rows, cols = 316, 317
            queries = random_strings(rng, rows * scale, 960, 1088, PROTEIN)
            candidates = random_strings(rng, cols, 960, 1088, PROTEIN)
            family = "smith_waterman" if name == "c3_sw" else "needleman_wunsch"
        gaps = dict(open=-4, extend=-4) if name == "c3_linear" else dict(open=-4, extend=-1)
        label = f"{name}: q{rows * scale} x c{cols} protein, {family} BLOSUM62 gaps {tuple(gaps.values())}"
        return label, family, dict(byte_to_class=byte_to_class, class_substitution_costs=table, **gaps), queries, candidates
    if name in ("s100_lev", "s100_nw", "s100_sw"):
        # The reference's other published table: ~100-byte strings, all-pairs tile (similarities/README.md:14-24).
        rows = cols = 2048
        alphabet = DNA if name == "s100_lev" else PROTEIN
        queries = random_strings(rng, rows * scale, 90, 110, alphabet)
        candidates = random_strings(rng, cols, 90, 110, alphabet)
        if name == "s100_lev":
            return f"{name}: q{rows * scale} x c{cols} DNA, len U[90,110], unit-cost Levenshtein", "levenshtein", dict(match=0, mismatch=1, open=1, extend=1), queries, candidates
        byte_to_class, table = scoring.blosum62()
        family = "needleman_wunsch" if name == "s100_nw" else "smith_waterman"
        return (f"{name}: q{rows * scale} x c{cols} protein, len U[90,110], {family} BLOSUM62 gap -4", family,
                dict(byte_to_class=byte_to_class, class_substitution_costs=table, open=-4, extend=-4), queries, candidates)
    raise SystemExit(f"unknown workload {name}")


def make_engine(family, kwargs, library_path=None, capabilities=None):
    import stringzilla_b200 as szb

    extra = dict(library_path=library_path, capabilities=capabilities)
    if family == "levenshtein":
        return szb.LevenshteinDistances(**kwargs, **extra)
    engine_type = szb.NeedlemanWunschScores if family == "needleman_wunsch" else szb.SmithWatermanScores
    return engine_type(**kwargs, **extra)


# ---------------------------------------------------------------------------------------------------------------------
# Clocks sampled DURING the timed region
# ---------------------------------------------------------------------------------------------------------------------


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, device_index: int):
        self.device_index, self.samples, self.process = device_index, [], None

    def __enter__(self):
        try:
            self.process = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.process = None
        return self

    def _read(self):
        for line in self.process.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                self.samples.append(parts)

    def __exit__(self, *exc):
        if self.process:
            time.sleep(0.15)
            self.process.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        clocks = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        if not clocks:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [c for c in clocks if c > 0.5 * max(clocks)] or clocks
        reasons = set()
        for s in self.samples:
            for name, flag in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": float(self.samples[0][2]), "reasons": sorted(reasons),
                "samples": len(clocks)}  # fmt: skip


# ---------------------------------------------------------------------------------------------------------------------
# The reference arm / CPU baseline: the reference's own CPU library (oracle/_ref), all host threads
# ---------------------------------------------------------------------------------------------------------------------


def allowed_cores() -> int:
    """Host threads this process may actually keep busy: the affinity mask, cut to the cgroup CPU quota when there is one
    (the reference's thread pool busy-waits between calls, so a pool wider than the quota throttles itself)."""
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = Path("/sys/fs/cgroup/cpu.max").read_text().split()[:2]  # cgroup v2
        if quota != "max":
            cores = min(cores, max(1, -(-int(quota) // int(period))))
    except (OSError, ValueError):
        try:  # cgroup v1
            quota = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read_text())
            period = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
            if quota > 0:
                cores = min(cores, max(1, -(-quota // period)))
        except (OSError, ValueError):
            pass
    return max(1, cores)


def time_reference(family, kwargs, queries, candidates, steps, warmup, budget_seconds=60.0):
    """Times the unmodified reference CPU build through its C ABI and keeps its best configuration: the serial and the
    SIMD capability masks (the survey found the SIMD engines not always faster when threaded), each with every core
    the library sees, with the cores this process is allowed to keep busy, and with a half and a quarter of that (its
    pool busy-waits, so oversubscription costs it dearly)."""
    import stringzilla_b200 as szb
    from stringzilla_b200 import _abi

    ref = ROOT / "oracle" / "_ref" / "libstringzillas_ref.so"
    if not ref.exists():
        return None
    q, c = szb.Tape(queries), szb.Tape(candidates)
    cells = int(q.lengths.sum()) * int(c.lengths.sum())
    full = szb.capabilities(ref)
    allowed = allowed_cores()
    core_counts = [0] + [n for n in dict.fromkeys((allowed, max(1, allowed // 2), max(1, allowed // 4)))]
    configurations = [(label, caps, n) for n in core_counts
                      for label, caps in (("simd", full), ("serial", _abi.CAP_SERIAL | _abi.CAP_PARALLEL))]
    best, seen_cores = None, set()
    for label, caps, requested in configurations:
        scope = szb.DeviceScope(cpu_cores=requested, library_path=ref)
        cores = scope.cpu_cores
        if (label, cores) in seen_cores:  # "all cores" and "allowed cores" coincide on an unrestricted host
            del scope
            continue
        seen_cores.add((label, cores))
        engine = make_engine(family, kwargs, library_path=ref, capabilities=caps)
        out = np.zeros((len(q), len(c)), dtype=engine._result_dtype)
        start = time.perf_counter()
        engine(q, c, device=scope, out=out)  # warm-up, also sizes the budget
        first = time.perf_counter() - start
        reps = max(1, min(steps, int(budget_seconds / len(configurations) / max(first, 1e-3))))
        for _ in range(max(0, min(warmup - 1, reps))):
            engine(q, c, device=scope, out=out)
        times = []
        for _ in range(reps):
            start = time.perf_counter()
            engine(q, c, device=scope, out=out)
            times.append(time.perf_counter() - start)
        mean = sum(times) / len(times)
        result = {"label": label, "seconds_per_step": mean, "gcups": cells / mean / 1e9, "steps": reps, "cores": cores,
                  "checksum": int(out.astype(np.uint64).sum() & 0xFFFFFFFF)}  # fmt: skip
        if best is None or result["gcups"] > best["gcups"]:
            best = result
        del engine, scope  # joins the pool before the next configuration is timed
    best["cells"] = cells
    return best


# ---------------------------------------------------------------------------------------------------------------------
# Our arm
# ---------------------------------------------------------------------------------------------------------------------


def ncu_dram_traffic(summary_name: str, kernel_substring: str):
    """DRAM bytes (read + write) per launch of a kernel, from the committed `ncu --set full` summary under profiles/
    (tools/ncu_summary.py output); None when the summary is absent. Sums the launches that make up one step."""
    path = ROOT / "profiles" / summary_name
    if not path.exists():
        return None
    total, inside, scale = 0.0, False, {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for line in path.read_text().splitlines():
        if line.startswith("## "):
            inside = kernel_substring in line
        elif inside and (line.startswith("dram__bytes_read.sum ") or line.startswith("dram__bytes_write.sum ")):
            parts = line.split()
            total += float(parts[1]) * scale.get(parts[2], 1)
    return int(total) if total else None


def dpx_peak_lane_ops_per_second():
    binary = ROOT / "tools" / "microbench" / "pipe_rates"
    if not binary.exists():
        return None
    try:
        line = subprocess.run([str(binary), "--dpx-peak"], capture_output=True, text=True, timeout=60).stdout.strip()
        return float(json.loads(line)["dpx_lane_ops_per_s"])
    except Exception:
        return None


def run_ours(args):
    import torch

    import stringzilla_b200 as szb
    from stringzilla_b200 import _abi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if args.gpus != world and rank == 0 and distributed:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if distributed:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    # Weak scaling (default): the tile grows with the GPU count. `--scaling strong` keeps BASELINE's tile and splits it.
    label, family, kwargs, queries, candidates = make_workload(args.workload, scale=world if args.scaling == "weak" else 1)
    q, c = szb.Tape(queries), szb.Tape(candidates)
    blocks = sharding.partition_rows(q.lengths, world)
    row_begin, row_end = blocks[rank]
    cells_total = int(q.lengths.sum()) * int(c.lengths.sum())
    cells_local = int(q.lengths[row_begin:row_end].sum()) * int(c.lengths.sum())

    engine = make_engine(family, kwargs)
    scope = szb.DeviceScope(gpu_device=local_rank)
    lib = _abi.bind()
    result_dtype = torch.int64

    # ---- leg 1: strings resident in HBM, scores left in HBM -------------------------------------------------------
    q_dev = torch.from_numpy(q.data).to(device)
    c_dev = torch.from_numpy(c.data).to(device)
    out_dev = torch.zeros((len(q), len(c)), dtype=result_dtype, device=device)
    tq = _abi.SequenceU64Tape(q_dev.data_ptr(), q.offsets.ctypes.data, len(q))  # offsets are host metadata
    tc = _abi.SequenceU64Tape(c_dev.data_ptr(), c.offsets.ctypes.data, len(c))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # larger than the 126 MB L2
    message = ct.c_char_p()

    def resident_step():
        status = lib.szs_b200_scores_u64tape_rows(engine._family_index, engine._handle, scope.handle, ct.byref(tq),
                                                  ct.byref(tc), row_begin, row_end, out_dev.data_ptr(), len(c),
                                                  ct.byref(message))
        _abi.check(status, message)
        return engine.last_call_stats

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_resident(step_function, current_engine):
        for _ in range(args.warmup):
            step_function()
        device_ms, kernel_ms, launches = [], [], 0
        barrier()
        with ClockSampler(local_rank) as sampler:
            for _ in range(args.steps):
                flush.fill_(1)  # evict the tile from L2 between timed iterations (outside the per-step event pair)
                torch.cuda.synchronize()
                stats = step_function()
                device_ms.append(stats.total_milliseconds)
                kernel_ms.append(stats.kernel_milliseconds)
                launches += int(stats.kernel_launches)
            barrier()
        return sum(device_ms) / 1e3, sum(kernel_ms) / 1e3, launches, sampler.summary(), current_engine.last_call_stats

    resident_seconds, kernel_seconds, launches, clocks, last_stats = timed_resident(resident_step, engine)

    # The int16 DPX wavefront tier on the same tile: unit-cost engines normally take the bit-parallel kernel, so the
    # tier the north star names is measured by pinning the engine to it (identical results, see tests).
    dpx_seconds = None
    if family == "levenshtein" and last_stats.pairs_bit_parallel:
        dpx_engine = make_engine(family, dict(kwargs, bit_parallel=False))

        def dpx_step():
            status = lib.szs_b200_scores_u64tape_rows(dpx_engine._family_index, dpx_engine._handle, scope.handle,
                                                      ct.byref(tq), ct.byref(tc), row_begin, row_end,
                                                      out_dev.data_ptr(), len(c), ct.byref(message))
            _abi.check(status, message)
            return dpx_engine.last_call_stats

        reference_scores = out_dev.clone()
        dpx_seconds = timed_resident(dpx_step, dpx_engine)
        assert torch.equal(reference_scores[row_begin:row_end], out_dev[row_begin:row_end]), "tiers disagree"

    # ---- leg 2: end to end through the C ABI with pinned host buffers ------------------------------------------------
    q_pin = torch.from_numpy(q.data).pin_memory()
    c_pin = torch.from_numpy(c.data).pin_memory()
    out_pin = torch.zeros((len(q), len(c)), dtype=result_dtype).pin_memory()
    hq = _abi.SequenceU64Tape(q_pin.data_ptr(), q.offsets.ctypes.data, len(q))
    hc = _abi.SequenceU64Tape(c_pin.data_ptr(), c.offsets.ctypes.data, len(c))

    def e2e_step():
        if not distributed:
            function = getattr(lib, f"szs_{engine._family}_u64tape")
            _abi.check(function(engine._handle, scope.handle, ct.byref(hq), ct.byref(hc), out_pin.data_ptr(), len(c),
                                ct.byref(message)), message)
            stats = engine.last_call_stats
            return int(stats.h2d_bytes), int(stats.d2h_bytes)
        # Rank 0 owns the host inputs: H2D there, one NCCL broadcast, row-sharded scoring, one NCCL gather, D2H.
        q_tape = q if rank == 0 else None
        c_tape = c if rank == 0 else None
        (qd, qo_host), (cd, co_host) = sharding.broadcast_tapes([q_tape, c_tape], src=0, device=device)
        ltq = _abi.SequenceU64Tape(qd.data_ptr(), qo_host.ctypes.data, len(qo_host) - 1)
        ltc = _abi.SequenceU64Tape(cd.data_ptr(), co_host.ctypes.data, len(co_host) - 1)
        _abi.check(lib.szs_b200_scores_u64tape_rows(engine._family_index, engine._handle, scope.handle, ct.byref(ltq),
                                                    ct.byref(ltc), row_begin, row_end, out_dev.data_ptr(), len(c),
                                                    ct.byref(message)), message)
        full = sharding.gather_rows(out_dev[row_begin:row_end], blocks, len(c), dst=0)
        if rank == 0:
            out_pin.copy_(full, non_blocking=False)
            return int(q.data.size + c.data.size + q.offsets.nbytes + c.offsets.nbytes), int(out_pin.numel() * 8)
        return 0, 0

    for _ in range(max(1, args.warmup)):
        h2d, d2h = e2e_step()
    barrier()
    start = time.perf_counter()
    for _ in range(args.steps):
        h2d, d2h = e2e_step()
    barrier()
    e2e_seconds = time.perf_counter() - start

    # ---- max over ranks -------------------------------------------------------------------------------------------
    dpx_pair = (dpx_seconds[0], dpx_seconds[1]) if dpx_seconds else (0.0, 0.0)
    times = torch.tensor([resident_seconds, kernel_seconds, e2e_seconds, *dpx_pair], dtype=torch.float64, device=device)
    if distributed:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    resident_seconds, kernel_seconds, e2e_seconds, dpx_resident_seconds, dpx_kernel_seconds = (float(t) for t in times)

    if rank == 0:
        value = cells_total * args.steps / resident_seconds / 1e9
        e2e_value = cells_total * args.steps / e2e_seconds / 1e9
        kernel_gcups = cells_total * args.steps / kernel_seconds / 1e9
        # Rooflines. Neither kernel is HBM-bound (~1e-5 algorithmic bytes per cell update): the bound is instruction
        # issue on the integer (ALU) pipe, whose rate — identical for DPX *.16x2, LOP3, SHF and IADD3 on this part —
        # is measured live by tools/microbench/pipe_rates --dpx-peak.
        lane_ops = dpx_peak_lane_ops_per_second()  # measured on rank 0's GPU; the peaks below scale it by the GPU count
        affine = kwargs.get("open") != kwargs.get("extend")
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        hbm_peak = json.loads(peaks_file.read_text())["hbm_gbs"] if peaks_file.exists() else 6650.0
        algorithmic_bytes = (int(q.data.size) + int(c.data.size)) + 8 * len(q) * len(c)

        def hbm_block(seconds):
            gbs = algorithmic_bytes * args.steps / seconds / 1e9
            return {"achieved_gbs": round(gbs, 2), "peak_gbs": hbm_peak, "frac": round(gbs / hbm_peak, 6),
                    "peak_source": "MEASURED_PEAKS.json" if peaks_file.exists() else "fallback"}

        def dpx_roofline(gcups, seconds):
            # 2 (linear) or 3 (affine) DPX-class instructions per packed pair of int16 cells in the
            # reference's formulation -> peak = lane_ops/s * 2 cells / k. `dpx_pipe_frac` counts THIS kernel's DPX
            # instructions (1 per pair for the shifted linear recurrence) against the measured DPX issue rate.
            k_reference = 3 if affine else 2
            k_ours = {"levenshtein": 3 if affine else 1, "needleman_wunsch": 3 if affine else 1,
                      "smith_waterman": 4.5 if affine else 2.5}[family]
            peak = lane_ops * 2 / k_reference / 1e9 * world
            return {"bound": "dpx-int16-issue", "achieved": round(gcups, 1), "peak": round(peak, 1), "unit": "GCUPS",
                    "frac": round(gcups / peak, 4),
                    "traffic": ncu_dram_traffic("r1_v2_wavefront_c2_weird.txt", "wavefront_kernel") if args.workload in ("c2", "c2_weird") else None,
                    "kernel": "szb::wavefront_kernel",
                    "dpx_instr_per_cell_pair": {"reference_formulation": k_reference, "this_kernel": k_ours},
                    "dpx_pipe_frac": round(gcups * 1e9 / 2 * k_ours / lane_ops / world, 4),
                    "lane_ops_per_s_measured": lane_ops, "hbm": hbm_block(seconds),
                    "kernel_ms_per_launch": round(seconds * 1e3 / args.steps, 4)}

        roofline, dpx_tier = None, None
        if lane_ops and last_stats.pairs_bit_parallel:
            # Bit-parallel kernel: 7 LOP3 + 2 SHF + 1 IADD3.X per 32-bit word = 10 ALU lane-ops per 32 DP cells.
            peak = lane_ops * 32 / 10 / 1e9 * world
            roofline = {"bound": "alu-issue (bit-parallel, 10 lane-ops per 32 cells)", "achieved": round(kernel_gcups, 1),
                        "peak": round(peak, 1), "unit": "GCUPS", "frac": round(kernel_gcups / peak, 4),
                        "traffic": ncu_dram_traffic("r1_v2_myers_c2.txt", "myers_kernel") if args.workload == "c2" else None,
                        "traffic_note": "DRAM read+write bytes per step from profiles/r1_v2_myers_c2.txt (ncu --set full); "
                                        f"algorithmic bytes per step: {algorithmic_bytes}",
                        "kernel": "szb::myers_kernel", "lane_ops_per_s_measured": lane_ops,
                        "hbm": hbm_block(kernel_seconds),
                        "kernel_ms_per_launch": round(kernel_seconds * 1e3 / args.steps, 4)}
        elif lane_ops:
            roofline = dpx_roofline(kernel_gcups, kernel_seconds)
        if dpx_seconds:
            dpx_kernel_gcups = cells_total * args.steps / dpx_kernel_seconds / 1e9
            dpx_tier = {"note": "same tile, engine pinned to the int16 DPX wavefront tier (szs_b200_set_bit_parallel(0))",
                        "value": round(cells_total * args.steps / dpx_resident_seconds / 1e9, 1), "unit": "GCUPS",
                        "ms_per_step": round(dpx_resident_seconds * 1e3 / args.steps, 4),
                        "roofline": dpx_roofline(dpx_kernel_gcups, dpx_kernel_seconds) if lane_ops else None}
        baseline = None
        if world == 1 and not args.skip_cpu_baseline:
            reference = time_reference(family, kwargs, queries, candidates, steps=3, warmup=1, budget_seconds=30)
            if reference:
                baseline = {"value": round(reference["gcups"], 2), "unit": "GCUPS", "cores": reference["cores"],
                            "kind": "reference", "sample": f"the full tile ({reference['cells']:.3e} cell updates) x "
                            f"{reference['steps']} steps, reference '{reference['label']}' capability mask, {reference['cores']} threads (best of all / allowed / half / quarter of the cores)"}  # fmt: skip
        line = {
            "metric": "GCUPS (giga cell-updates/s), batched Levenshtein at |s|~1024" if args.workload == "c2" else f"GCUPS, {family}",
            "value": round(value, 1), "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(resident_seconds * 1e3 / args.steps, 4), "higher_is_better": True, "scaling": args.scaling,
            # BASELINE.md §1 publishes one figure for this metric: the reference's own CUDA path on an H100 at 1000 bytes.
            "vs_baseline": round(value / world / 5910.8, 3) if args.workload == "c2" else None,
            "vs_baseline_note": ("per-GPU value / 5910.8 GCUPS = the reference's published H100 figure for 1000-byte unit-cost "
                                 "Levenshtein (include/stringzillas/similarities/README.md:35); other hardware. Same box: "
                                 "profiles/r1_reference_cuda_same_box.jsonl") if args.workload == "c2" else None,
            "dtype": ("u32 bit-vectors (bit-parallel)" if last_stats.pairs_bit_parallel else
                      "u16x2 (packed int16 DPX cells)" if last_stats.pairs_u16 else "s32"),
            "data": "synthetic", "impl": "ours",
            "config": {"workload": label, "tile_rows_per_gpu": len(q) // world, "l2": "flushed between timed iterations (256 MB fill)",
                       "partition": f"query rows over {world} rank(s), balanced by total query length"},
            "e2e": {"value": round(e2e_value, 1), "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_seconds * 1e3 / args.steps, 4)},
            "gpu_launches": launches, "kernel_only_gcups": round(kernel_gcups, 1),
            "clocks": clocks, "roofline": roofline, "dpx_tier": dpx_tier, "cpu_baseline": baseline,
            "cells_per_step": cells_total,
        }  # fmt: skip
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle/_ref), all host threads, same
    config. Under torchrun only rank 0 runs; the other ranks exit 0 without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    label, family, kwargs, queries, candidates = make_workload(args.workload, scale=1)
    reference = time_reference(family, kwargs, queries, candidates, steps=args.steps, warmup=args.warmup, budget_seconds=120)
    if reference is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libstringzillas_ref.so is not built"}))
        return
    value = round(reference["gcups"], 2)
    sample = (f"one tile per step ({reference['cells']:.3e} cell updates), {reference['steps']} timed steps, reference "
              f"'{reference['label']}' capability mask, {reference['cores']} threads (best of all / allowed / half / quarter of the "
              f"host's cores; the reference's pool busy-waits, so oversubscribing a CPU quota throttles it)")
    print(json.dumps({
        "metric": "GCUPS (giga cell-updates/s), batched Levenshtein at |s|~1024" if args.workload == "c2" else f"GCUPS, {family}",
        "value": value, "unit": "GCUPS", "n_gpus": world, "steps": reference["steps"], "warmup": args.warmup,
        "ms_per_step": round(reference["seconds_per_step"] * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 bit-parallel / u8..u64 DP cells (reference CPU tiers)", "data": "synthetic",
        "impl": "reference", "config": {"workload": label},
        "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": reference["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))  # fmt: skip


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=200)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", choices=["ours", "reference"], default="ours")
    parser.add_argument("--workload", default="c2")
    parser.add_argument("--skip-cpu-baseline", action="store_true")
    parser.add_argument("--scaling", choices=["weak", "strong"], default="weak")
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
