#!/usr/bin/env python3
"""bench.py — cell-updates/s of the batched string-similarity hot path, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

A *step* is one pass of the hot path over one synthetic tile: a single `szs_*_u64tape` cross-product call.
The default workload is BASELINE.json `configs[1]`: ~100k (query, candidate) pairs of DNA strings with lengths
uniform on [960, 1088], unit-cost Levenshtein — the tile `q316 x c317` of SURVEY.md §8(d).

Numbers on the line
  value        GCUPS with the strings already resident in HBM and the score matrix left in HBM; time = CUDA events
               recorded HERE on the library's own stream around every call (staging D2D + scheduling + kernels + D2D),
               max over ranks.
  e2e          the same metric through the C-ABI call with pinned HOST buffers: H2D of the strings and D2H of the
               score matrix inside the timed region (wall clock around the synchronous calls, max over ranks).
  parity       the matrices of the timed runs equal the reference CPU library's on the same tile (oracle/_ref as the
               checker, like bench/similarities.cuh:410-423 of the reference); a mismatch exits non-zero.
  roofline     the dominant kernel against the integer-issue roofline measured live by tools/microbench/pipe_rates.
  workloads    every other BASELINE config (c1, c2_weird, c3, c3_sw, c4, c5), each with value / e2e / roofline / parity.
  reference_cuda_same_box  the reference's own CUDA library rebuilt for sm_100, same tiles, same GPU, matrices compared.
  cpu_baseline the reference's own CPU library (oracle/_ref) on this box's host cores, same tile.
N > 1 (launched by torch.distributed.run): the headline is weak-scaled c2 — the tile grows to 316*N query rows, rank r
scores its row block (balanced by total query length); the e2e leg is `szs_b200_comm_scores_u64tape`: ONE NCCL broadcast
of the packed inputs from rank 0 and ONE NCCL gather of the score slabs, issued from C. `workloads` then holds c4 and c5
STRONG-scaled (BASELINE configs[3], [4]), gathered matrices checked on rank 0.
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
# Parity inside the benchmark: the timed GPU matrices against the reference's own CPU library on the same tile
# (what the reference's bench does on every batch, bench/similarities.cuh:410-423)
# ---------------------------------------------------------------------------------------------------------------------

PARITY_SAMPLE_ROWS = {"c4": 24, "c5": 8}  # tiles whose full matrix would take the CPU library minutes: sampled rows


def reference_rows(family, kwargs, queries, candidates, rows):
    """Scores the given query rows against every candidate with the UNMODIFIED reference CPU library (oracle/_ref), all
    host threads the process may use. Checker only: never timed here, never on the product path."""
    import stringzilla_b200 as szb

    ref = ROOT / "oracle" / "_ref" / "libstringzillas_ref.so"
    if not ref.exists():
        return None
    scope = szb.DeviceScope(cpu_cores=allowed_cores(), library_path=ref)
    engine = make_engine(family, kwargs, library_path=ref, capabilities=szb.capabilities(ref))
    return engine([queries[i] for i in rows], candidates, device=scope)


def check_parity(name, family, kwargs, queries, candidates, ours):
    """`ours` is the full (queries x candidates) matrix of the timed GPU run, on the host. Returns the parity record."""
    sample = PARITY_SAMPLE_ROWS.get(name)
    n = len(queries)
    if sample and n > sample:
        rows = sorted(set(np.linspace(0, n - 1, sample).astype(int).tolist()))
        against = f"{len(rows)} sampled query rows x all {len(candidates)} candidates"
    else:
        rows, against = list(range(n)), "the full matrix"
    want = reference_rows(family, kwargs, queries, candidates, rows)
    if want is None:
        return {"parity": None, "against": "oracle/_ref/libstringzillas_ref.so is not built"}
    equal = bool(np.array_equal(np.asarray(ours)[rows].astype(np.int64), want.astype(np.int64)))
    return {"parity": equal, "against": f"reference CPU library (oracle/_ref), {against}",
            "cells_compared": int(len(rows) * len(candidates))}


# ---------------------------------------------------------------------------------------------------------------------
# The reference's own CUDA library rebuilt for sm_100 (oracle/build_ref_cuda.sh), timed on the same box
# ---------------------------------------------------------------------------------------------------------------------


class ManagedTape:
    """A u64 tape whose blob and offsets live in `szs_unified_alloc` memory, as the reference's GPU engines require
    (include/stringzillas/similarities/cuda.cuh:4268-4272) and as its Python binding hands strings over."""

    def __init__(self, lib, tape):
        from stringzilla_b200 import _abi

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


def time_reference_cuda(name, ours_matrix, repeats=5):
    """One workload through `oracle/_ref/libstringzillas_ref_cuda.so` (the reference's CUDA + Kepler + Hopper tiers
    compiled unmodified with -arch=sm_100) on GPU 0: wall clock around the synchronous C-ABI call with managed-memory
    tapes, best of `repeats` after one warm-up, and whether its matrix equals ours."""
    import stringzilla_b200 as szb
    from stringzilla_b200 import _abi

    ref = ROOT / "oracle" / "_ref" / "libstringzillas_ref_cuda.so"
    if not ref.exists():
        return {"unavailable": "oracle/_ref/libstringzillas_ref_cuda.so is not built"}
    try:
        lib = _abi.bind(ref)
        caps = int(lib.szs_capabilities()) & _abi.CAPS_CUDA  # GPU tiers only: the init ladder prefers CPU SIMD tiers
        if not caps:
            return {"unavailable": "the reference CUDA library reports no GPU capability on this box"}
        scope = szb.DeviceScope(gpu_device=0, library_path=ref)
        label, family, kwargs, queries, candidates = make_workload(name)
        q, c = szb.Tape(queries), szb.Tape(candidates)
        cells = int(q.lengths.sum()) * int(c.lengths.sum())
        engine = make_engine(family, kwargs, library_path=ref, capabilities=caps)
        mq, mc = ManagedTape(lib, q), ManagedTape(lib, c)
        out = np.zeros((len(q), len(c)), dtype=engine._result_dtype)
        function = getattr(lib, f"szs_{engine._family}_u64tape")
        message, times = ct.c_char_p(), []
        for _ in range(repeats + 1):
            start = time.perf_counter()
            status = function(engine._handle, scope.handle, mq.pointer(), mc.pointer(), out.ctypes.data, len(c), ct.byref(message))
            times.append(time.perf_counter() - start)
            _abi.check(status, message)
        best = min(times[1:])
        return {"gcups": round(cells / best / 1e9, 1), "ms_per_call": round(best * 1e3, 3), "capabilities": hex(caps),
                "timing": f"wall clock around the synchronous call, managed-memory tapes, best of {repeats}",
                "matches_ours": bool(np.array_equal(out.astype(np.int64), np.asarray(ours_matrix).astype(np.int64)))}
    except Exception as error:  # the comparison arm must never take the benchmark down
        return {"unavailable": f"{type(error).__name__}: {error}"}


# ---------------------------------------------------------------------------------------------------------------------
# Our arm
# ---------------------------------------------------------------------------------------------------------------------


def ncu_dram_traffic(summary_names, kernel_substring: str):
    """DRAM bytes (read + write) per step of a kernel family, from the newest committed `ncu --set full` summary under
    profiles/ (tools/ncu_summary.py output); None when no summary is there. Sums the launches that make up one step."""
    for summary_name in summary_names:
        path = ROOT / "profiles" / summary_name
        if not path.exists():
            continue
        total, inside, scale = 0.0, False, {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for line in path.read_text().splitlines():
            if line.startswith("## "):
                inside = kernel_substring in line
            elif inside and (line.startswith("dram__bytes_read.sum ") or line.startswith("dram__bytes_write.sum ")):
                parts = line.split()
                total += float(parts[1]) * scale.get(parts[2], 1)
        if total:
            return int(total), summary_name
    return None, None


def dpx_peak_lane_ops_per_second():
    binary = ROOT / "tools" / "microbench" / "pipe_rates"
    if not binary.exists():
        return None
    try:
        line = subprocess.run([str(binary), "--dpx-peak"], capture_output=True, text=True, timeout=60).stdout.strip()
        return float(json.loads(line)["dpx_lane_ops_per_s"])
    except Exception:
        return None


TRAFFIC_SUMMARIES = {
    "myers": (["r2_myers_c2.txt", "r1_v2_myers_c2.txt"], "myers_kernel", ("c2",)),
    "wavefront": (["r2_wavefront_c2_weird.txt", "r1_v2_wavefront_c2_weird.txt"], "wavefront_kernel", ("c2", "c2_weird")),
}


class Bench:
    """One process = one rank = one GPU. Holds the library handles and measures one workload at a time."""

    def __init__(self, args):
        import torch

        import stringzilla_b200 as szb
        from stringzilla_b200 import _abi

        self.args, self.torch, self.szb, self.abi = args, torch, szb, _abi
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.distributed = self.world > 1
        if args.gpus != self.world and self.rank == 0 and self.distributed:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE={self.world}", file=sys.stderr)
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        self.lib = _abi.bind()
        self.scope = szb.DeviceScope(gpu_device=self.local_rank)
        self.stream = torch.cuda.ExternalStream(self.scope.stream, device=self.device)  # the library's own stream
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.device)  # larger than the 126 MB L2
        self.communicator = None
        if self.distributed:
            import torch.distributed as dist

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            import datetime

            # A short watchdog: a rank that fails must not leave its peers spinning in a collective for ten minutes.
            dist.init_process_group("nccl", device_id=self.device, timeout=datetime.timedelta(seconds=180))
            self.dist = dist
            # torch.distributed is plumbing only: it carries the 128-byte NCCL identifier to the other ranks; the data
            # path of the e2e leg is the library's own communicator (one broadcast + one gather issued from C).
            holder = [szb.Communicator.unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(holder, src=0)
            self.communicator = szb.Communicator(holder[0], self.world, self.rank, self.local_rank)

    # -- helpers ------------------------------------------------------------------------------------------------------

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.distributed:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        tensor = self.torch.tensor(values, dtype=self.torch.float64, device=self.device)
        if self.distributed:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in tensor]

    def min_over_ranks(self, value):
        tensor = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.device)
        if self.distributed:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.MIN)
        return float(tensor[0])

    def partition(self, lengths, symmetric=False):
        bounds = np.zeros(self.world + 1, dtype=np.uint64)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint64)
        self.lib.szs_b200_partition_rows(lengths.ctypes.data, lengths.size, self.world, int(symmetric), bounds.ctypes.data)
        return [(int(bounds[r]), int(bounds[r + 1])) for r in range(self.world)]

    # -- one workload -------------------------------------------------------------------------------------------------

    def measure(self, name, steps, warmup, scaling, with_dpx_tier=False, with_clocks=False, budget_seconds=None):
        """Returns (record, extras) on rank 0 and (None, None) elsewhere. `record` carries value / e2e / parity."""
        torch, szb, _abi, lib = self.torch, self.szb, self.abi, self.lib
        world, rank = self.world, self.rank
        label, family, kwargs, queries, candidates = make_workload(name, scale=world if scaling == "weak" else 1)
        q, c = szb.Tape(queries), szb.Tape(candidates)
        blocks = self.partition(q.lengths)
        row_begin, row_end = blocks[rank]
        cells_total = int(q.lengths.sum()) * int(c.lengths.sum())
        engine = make_engine(family, kwargs)
        message = ct.c_char_p()

        # ---- leg 1: strings resident in HBM, scores left in HBM. Timed HERE with CUDA events on the library's stream.
        q_dev = torch.from_numpy(q.data).to(self.device)
        c_dev = torch.from_numpy(c.data).to(self.device)
        out_dev = torch.zeros((len(q), len(c)), dtype=torch.int64, device=self.device)
        tq = _abi.SequenceU64Tape(q_dev.data_ptr(), q.offsets.ctypes.data, len(q))  # offsets are host metadata
        tc = _abi.SequenceU64Tape(c_dev.data_ptr(), c.offsets.ctypes.data, len(c))

        def resident_step(current):
            _abi.check(lib.szs_b200_scores_u64tape_rows(current._family_index, current._handle, self.scope.handle,
                                                        ct.byref(tq), ct.byref(tc), row_begin, row_end,
                                                        out_dev.data_ptr(), len(c), ct.byref(message)), message)
            return current.last_call_stats

        def timed_resident(current, reps):
            first_call = time.perf_counter()
            resident_step(current)
            first_call = time.perf_counter() - first_call
            if budget_seconds:
                reps = max(2, min(reps, int(budget_seconds / max(first_call, 1e-4))))
                reps = int(self.min_over_ranks(reps))  # every rank must enter the same number of collective steps
            for _ in range(max(0, warmup - 1)):
                resident_step(current)
            events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
            kernel_ms, launches = 0.0, 0
            self.barrier()
            sampler = ClockSampler(self.local_rank) if with_clocks else None
            if sampler:
                sampler.__enter__()
            for begin, end in events:
                self.flush.fill_(1)  # evict the tile from L2 between timed iterations (outside the event pair)
                torch.cuda.synchronize()
                begin.record(self.stream)
                stats = resident_step(current)
                end.record(self.stream)
                kernel_ms += stats.kernel_milliseconds
                launches += int(stats.kernel_launches)
            self.barrier()
            if sampler:
                sampler.__exit__(None, None, None)
            device_ms = sum(begin.elapsed_time(end) for begin, end in events)
            return {"seconds": device_ms / 1e3, "kernel_seconds": kernel_ms / 1e3, "launches": launches, "reps": reps,
                    "clocks": sampler.summary() if sampler else None, "stats": current.last_call_stats}

        resident = timed_resident(engine, steps)
        resident_rows = out_dev[row_begin:row_end].cpu().numpy() if rank == 0 else None

        dpx = None
        if with_dpx_tier and family == "levenshtein" and resident["stats"].pairs_bit_parallel:
            # The int16 DPX wavefront tier on the same tile: unit-cost engines normally take the bit-parallel kernel,
            # so the tier the north star names is measured by pinning the engine to it.
            dpx_engine = make_engine(family, dict(kwargs, bit_parallel=False))
            reference_scores = out_dev[row_begin:row_end].clone()
            dpx = timed_resident(dpx_engine, steps)
            dpx["equal_to_bit_parallel"] = bool(torch.equal(reference_scores, out_dev[row_begin:row_end]))

        # ---- leg 2: end to end through the C ABI with pinned HOST buffers on rank 0 -----------------------------------
        if rank == 0:
            q_pin = torch.from_numpy(q.data).pin_memory()
            c_pin = torch.from_numpy(c.data).pin_memory()
            out_pin = torch.zeros((len(q), len(c)), dtype=torch.int64).pin_memory()
            hq = _abi.SequenceU64Tape(q_pin.data_ptr(), q.offsets.ctypes.data, len(q))
            hc = _abi.SequenceU64Tape(c_pin.data_ptr(), c.offsets.ctypes.data, len(c))
        function = getattr(lib, f"szs_{engine._family}_u64tape")

        def e2e_step():
            if not self.distributed:
                _abi.check(function(engine._handle, self.scope.handle, ct.byref(hq), ct.byref(hc), out_pin.data_ptr(), len(c),
                                    ct.byref(message)), message)
                return
            # One process per GPU: rank 0 owns the host buffers; ONE NCCL broadcast of the packed inputs, row-sharded
            # scoring, ONE NCCL gather of the score slabs, D2H on rank 0 — all issued from C (szs_b200_comm_*).
            if rank == 0:
                status = lib.szs_b200_comm_scores_u64tape(engine._family_index, engine._handle, self.scope.handle,
                                                          self.communicator.handle, ct.byref(hq), ct.byref(hc),
                                                          out_pin.data_ptr(), len(c), ct.byref(message))
            else:
                status = lib.szs_b200_comm_scores_u64tape(engine._family_index, engine._handle, self.scope.handle,
                                                          self.communicator.handle, None, None, None, 0, ct.byref(message))
            _abi.check(status, message)

        for _ in range(max(1, min(warmup, 2))):
            e2e_step()
        e2e_reps = resident["reps"]
        self.barrier()
        start = time.perf_counter()
        for _ in range(e2e_reps):
            e2e_step()
        self.barrier()
        e2e_seconds = time.perf_counter() - start

        # ---- a managed-memory call, the way the reference's Python binding hands strings over (N = 1 only) ------------
        managed_seconds = None
        if not self.distributed and name in ("c2", "c3"):
            mq, mc = ManagedTape(lib, q), ManagedTape(lib, c)
            out_managed = lib.szs_unified_alloc(len(q) * len(c) * 8)
            for timed in (False, True):
                start = time.perf_counter()
                for _ in range(3):
                    _abi.check(function(engine._handle, self.scope.handle, mq.pointer(), mc.pointer(), out_managed, len(c),
                                        ct.byref(message)), message)
                if timed:
                    managed_seconds = (time.perf_counter() - start) / 3
            lib.szs_unified_free(out_managed, len(q) * len(c) * 8)

        seconds, kernel_seconds, e2e_seconds, dpx_seconds, dpx_kernel_seconds = self.max_over_ranks(
            [resident["seconds"] / resident["reps"], resident["kernel_seconds"] / resident["reps"], e2e_seconds / e2e_reps,
             (dpx["seconds"] / dpx["reps"]) if dpx else 0.0, (dpx["kernel_seconds"] / dpx["reps"]) if dpx else 0.0])
        if rank != 0:
            return None, None

        # ---- parity gate: the matrices this run produced against the reference's CPU library --------------------------
        e2e_matrix = out_pin.numpy()
        parity = check_parity(name, family, kwargs, queries, candidates, e2e_matrix)
        if parity["parity"] is not None:
            parity["parity"] = bool(parity["parity"] and np.array_equal(resident_rows, e2e_matrix[row_begin:row_end]))
            parity["against"] += "; HBM-resident leg == host-buffer leg"
        if dpx is not None:
            parity["parity"] = bool(parity["parity"] and dpx["equal_to_bit_parallel"]) if parity["parity"] is not None else None
        stats = resident["stats"]
        affine = kwargs.get("open") != kwargs.get("extend")
        record = {
            "workload": label, "scaling": scaling if self.distributed else "single GPU",
            "value": round(cells_total / seconds / 1e9, 1), "unit": "GCUPS", "ms_per_step": round(seconds * 1e3, 4),
            "kernel_only_gcups": round(cells_total / kernel_seconds / 1e9, 1), "steps": resident["reps"],
            "e2e": {"value": round(cells_total / e2e_seconds / 1e9, 1), "unit": "GCUPS", "ms_per_step": round(e2e_seconds * 1e3, 4),
                    "h2d_bytes_per_step": int(q.data.size + c.data.size + q.offsets.nbytes + c.offsets.nbytes),
                    "d2h_bytes_per_step": int(len(q) * len(c) * 8),
                    "path": ("szs_*_u64tape, pinned host buffers" if not self.distributed else
                             "szs_b200_comm_scores_u64tape: pinned host buffers on rank 0, one NCCL broadcast + one NCCL gather")},
            "kernel": ("szb::myers_kernel (bit-parallel)" if stats.pairs_bit_parallel else
                       "szb::wavefront_kernel (packed int16 DPX)" if stats.pairs_u16 else "szb::wavefront_kernel (s32 DPX)"),
            "gpu_launches_per_step": resident["launches"] // max(1, resident["reps"]),
            "cells_per_step": cells_total, **parity,
        }
        if managed_seconds:
            record["e2e_managed"] = {"value": round(cells_total / managed_seconds / 1e9, 1), "unit": "GCUPS",
                                     "ms_per_step": round(managed_seconds * 1e3, 4),
                                     "path": "szs_*_u64tape, inputs and results in szs_unified_alloc (managed) memory"}
        extras = {"kernel_seconds": kernel_seconds, "seconds": seconds, "e2e_seconds": e2e_seconds, "stats": stats, "affine": affine,
                  "family": family, "kwargs": kwargs, "queries": queries, "candidates": candidates, "q": q, "c": c,
                  "clocks": resident["clocks"], "launches": resident["launches"], "reps": resident["reps"], "dpx": dpx,
                  "dpx_seconds": dpx_seconds, "dpx_kernel_seconds": dpx_kernel_seconds, "cells": cells_total,
                  "matrix": e2e_matrix, "label": label}
        return record, extras


def rooflines(lane_ops, world):
    """Issue rooflines of the two kernel families, from the integer-pipe rate measured live on this GPU
    (tools/microbench/pipe_rates --dpx-peak; identical for DPX *.16x2, LOP3, SHF and IADD3 on this part)."""

    def wavefront(family, affine, gcups, kernel_seconds, workload_name, hbm):
        # 2 (linear) or 3 (affine) DPX-class instructions per packed pair of int16 cells in the
        # reference's formulation -> peak = lane_ops/s * 2 cells / k. `dpx_pipe_frac` counts THIS kernel's DPX
        # instructions per pair against the measured DPX issue rate.
        k_reference = 3 if affine else 2
        k_ours = {"levenshtein": 3 if affine else 1, "needleman_wunsch": 3 if affine else 1,
                  "smith_waterman": 3.5 if affine else 2.5}[family]
        peak = lane_ops * 2 / k_reference / 1e9 * world
        names, kernel, workloads = TRAFFIC_SUMMARIES["wavefront"]
        traffic, source = ncu_dram_traffic(names, kernel) if workload_name in workloads else (None, None)
        return {"bound": "dpx-int16-issue", "achieved": round(gcups, 1), "peak": round(peak, 1), "unit": "GCUPS",
                "frac": round(gcups / peak, 4), "traffic": traffic, "traffic_source": source, "kernel": "szb::wavefront_kernel",
                "dpx_instr_per_cell_pair": {"reference_formulation": k_reference, "this_kernel": k_ours},
                "dpx_pipe_frac": round(gcups * 1e9 / 2 * k_ours / lane_ops / world, 4),
                "lane_ops_per_s_measured": lane_ops, "hbm": hbm, "kernel_ms_per_step": round(kernel_seconds * 1e3, 4)}

    def myers(gcups, kernel_seconds, workload_name, hbm, algorithmic_bytes):
        # Bit-parallel kernel: 7 LOP3 + 2 shifts + 1 IADD3.X per 32-bit word = 10 integer lane-ops per 32 DP cells.
        peak = lane_ops * 32 / 10 / 1e9 * world
        names, kernel, workloads = TRAFFIC_SUMMARIES["myers"]
        traffic, source = ncu_dram_traffic(names, kernel) if workload_name in workloads else (None, None)
        return {"bound": "alu-issue (bit-parallel, 10 lane-ops per 32 cells)", "achieved": round(gcups, 1),
                "peak": round(peak, 1), "unit": "GCUPS", "frac": round(gcups / peak, 4), "traffic": traffic,
                "traffic_note": f"DRAM read+write bytes per step from profiles/{source} (ncu --set full); algorithmic bytes per step: {algorithmic_bytes}",
                "kernel": "szb::myers_kernel", "lane_ops_per_s_measured": lane_ops, "hbm": hbm,
                "kernel_ms_per_step": round(kernel_seconds * 1e3, 4)}

    return wavefront, myers


def run_ours(args):
    bench = Bench(args)
    world, rank = bench.world, bench.rank
    main_name = args.workload
    record, extras = bench.measure(main_name, args.steps, args.warmup, args.scaling, with_dpx_tier=True, with_clocks=True)

    # The other BASELINE configs, each a parity-gated measurement of its own (N > 1: the configs BASELINE.json names as
    # multi-GPU jobs, strong-scaled: the tile stays, the rows are split).
    if args.workloads is not None:
        names = [n for n in args.workloads.split(",") if n]
    elif bench.distributed:
        names = ["c4", "c5"]
    else:
        names = ["c1", "c2_weird", "c3", "c3_sw", "c4", "c5"]
    others = {}
    for name in names:
        if name == main_name:
            continue
        others[name] = bench.measure(name, args.steps, 2, "strong", budget_seconds=args.workload_budget)

    if rank == 0:
        lane_ops = dpx_peak_lane_ops_per_second()  # measured on rank 0's GPU; the peaks scale it by the GPU count
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        hbm_peak = json.loads(peaks_file.read_text())["hbm_gbs"] if peaks_file.exists() else 6650.0
        wavefront_roofline, myers_roofline = rooflines(lane_ops, world) if lane_ops else (None, None)

        def roofline_of(name, rec, ext):
            if not lane_ops:
                return None
            algorithmic_bytes = (int(ext["q"].data.size) + int(ext["c"].data.size)) + 8 * len(ext["q"]) * len(ext["c"])
            gbs = algorithmic_bytes / ext["kernel_seconds"] / 1e9
            hbm = {"achieved_gbs": round(gbs, 2), "peak_gbs": hbm_peak, "frac": round(gbs / hbm_peak, 6),
                   "peak_source": "MEASURED_PEAKS.json" if peaks_file.exists() else "fallback"}
            kernel_gcups = ext["cells"] / ext["kernel_seconds"] / 1e9
            if ext["stats"].pairs_bit_parallel:
                return myers_roofline(kernel_gcups, ext["kernel_seconds"], name, hbm, algorithmic_bytes)
            return wavefront_roofline(ext["family"], ext["affine"], kernel_gcups, ext["kernel_seconds"], name, hbm)

        roofline = roofline_of(main_name, record, extras)
        dpx_tier = None
        if extras["dpx"]:
            dpx_kernel_gcups = extras["cells"] / extras["dpx_kernel_seconds"] / 1e9
            dpx_tier = {"note": "same tile, engine pinned to the int16 DPX wavefront tier (szs_b200_set_bit_parallel(0)); "
                                "its matrix equals the bit-parallel one",
                        "value": round(extras["cells"] / extras["dpx_seconds"] / 1e9, 1), "unit": "GCUPS",
                        "ms_per_step": round(extras["dpx_seconds"] * 1e3, 4),
                        "roofline": wavefront_roofline("levenshtein", False, dpx_kernel_gcups, extras["dpx_kernel_seconds"],
                                                       main_name, None) if lane_ops else None}
        workloads = {}
        for name, (rec, ext) in others.items():
            rec["roofline_frac"] = (roofline_of(name, rec, ext) or {}).get("frac")
            rec["roofline_bound"] = (roofline_of(name, rec, ext) or {}).get("bound")
            workloads[name] = rec

        # The reference's own CUDA library on this very GPU (N = 1): same tiles, matrices compared with ours.
        reference_cuda = None
        if world == 1 and not args.skip_reference_cuda:
            reference_cuda = {}
            matrices = {main_name: extras["matrix"], **{n: e["matrix"] for n, (r, e) in others.items()}}
            for name in ("c2", "c2_weird", "c3", "c3_sw"):
                if name in matrices:
                    entry = time_reference_cuda(name, matrices[name])
                    if "gcups" in entry:
                        ours = record["value"] if name == main_name else workloads[name]["value"]
                        entry["ours_over_reference_cuda"] = round(ours / entry["gcups"], 2)
                    reference_cuda[name] = entry

        baseline = None
        if world == 1 and not args.skip_cpu_baseline:
            reference = time_reference(extras["family"], extras["kwargs"], extras["queries"], extras["candidates"], steps=3,
                                       warmup=1, budget_seconds=30)
            if reference:
                baseline = {"value": round(reference["gcups"], 2), "unit": "GCUPS", "cores": reference["cores"],
                            "kind": "reference", "sample": f"the full tile ({reference['cells']:.3e} cell updates) x "
                            f"{reference['steps']} steps, reference '{reference['label']}' capability mask, {reference['cores']} threads (best of all / allowed / half / quarter of the cores)"}  # fmt: skip
        stats = extras["stats"]
        parities = [record["parity"]] + [w["parity"] for w in workloads.values()]
        line = {
            "metric": "GCUPS (giga cell-updates/s), batched Levenshtein at |s|~1024" if main_name == "c2" else f"GCUPS, {extras['family']}",
            "value": record["value"], "unit": "GCUPS", "n_gpus": world, "steps": extras["reps"], "warmup": args.warmup,
            "ms_per_step": record["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
            # BASELINE.md §1 publishes one figure for this metric: the reference's own CUDA path on an H100 at 1000 bytes.
            "vs_baseline": round(record["value"] / world / 5910.8, 3) if main_name == "c2" else None,
            "vs_baseline_note": ("per-GPU value / 5910.8 GCUPS = the reference's published H100 figure for 1000-byte unit-cost "
                                 "Levenshtein (include/stringzillas/similarities/README.md:35); other hardware. Same box: "
                                 "`reference_cuda_same_box` below") if main_name == "c2" else None,
            "dtype": ("u32 bit-vectors (bit-parallel)" if stats.pairs_bit_parallel else
                      "u16x2 (packed int16 DPX cells)" if stats.pairs_u16 else "s32"),
            "data": "synthetic", "impl": "ours",
            "config": {"workload": extras["label"], "tile_rows_per_gpu": len(extras["q"]) // world,
                       "l2": "flushed between timed iterations (256 MB fill)",
                       "timing": "value: CUDA events recorded by bench.py on the library's own stream around each call; e2e: wall clock",
                       "partition": f"query rows over {world} rank(s), balanced by total query length (szs_b200_partition_rows)"},
            "e2e": record["e2e"], "e2e_managed": record.get("e2e_managed"),
            "gpu_launches": extras["launches"], "kernel_only_gcups": record["kernel_only_gcups"],
            "parity": record["parity"], "parity_against": record["against"], "parity_all_workloads": all(p is not False for p in parities),
            "clocks": extras["clocks"], "roofline": roofline, "dpx_tier": dpx_tier, "cpu_baseline": baseline,
            "workloads": workloads, "reference_cuda_same_box": reference_cuda,
            "cells_per_step": extras["cells"],
        }  # fmt: skip
        print(json.dumps(line))
        sys.stdout.flush()
        failed = [name for name, parity in zip([main_name] + list(workloads), parities) if parity is False]
    else:
        failed = []
    if bench.distributed:
        bench.dist.barrier()
        bench.dist.destroy_process_group()
    if failed:
        print(f"PARITY FAILURE against the reference library on: {failed}", file=sys.stderr)
        sys.exit(1)


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle/_ref), all host threads, same
    config. Under torchrun only rank 0 runs; the other ranks exit 0 without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # The same tile our arm scores at this GPU count: weak scaling grows the query rows with the number of ranks.
    label, family, kwargs, queries, candidates = make_workload(args.workload, scale=world if args.scaling == "weak" else 1)
    reference = time_reference(family, kwargs, queries, candidates, steps=args.steps, warmup=args.warmup, budget_seconds=80)
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
        "ms_per_step": round(reference["seconds_per_step"] * 1e3, 3), "higher_is_better": True, "scaling": args.scaling,
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
    parser.add_argument("--skip-reference-cuda", action="store_true")
    parser.add_argument("--workloads", default=None,
                        help="comma-separated extra workloads for the `workloads` map ('' for none); default: every other "
                             "BASELINE config at N=1, c4 and c5 (strong-scaled) at N>1")
    parser.add_argument("--workload-budget", type=float, default=1.5,
                        help="seconds of timed steps per leg of each extra workload (at least 2 steps)")
    parser.add_argument("--scaling", choices=["weak", "strong"], default="weak")
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
