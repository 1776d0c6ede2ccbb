#!/usr/bin/env python3
"""Prints the markdown rows of DESIGN.md §7 from bench.py JSON lines (one file per GPU count):
    python tools/measured_table.py profiles/r2_bench_1gpu.json [profiles/r2_bench_2gpu.json ...]"""
import json
import sys


def load(path):
    lines = [line for line in open(path).read().splitlines() if line.startswith("{")]
    return json.loads(lines[-1])


def tcups(gcups):
    return f"{gcups / 1e3:.2f}"


for path in sys.argv[1:]:
    d = load(path)
    n = d["n_gpus"]
    print(f"\n### {path}: {n} GPU(s), clocks {d['clocks']}\n")
    print("| workload | kernel | value (HBM-resident) TCUPS | ms/step | e2e (host buffers) TCUPS | roofline frac | parity vs reference CPU library | reference CUDA same box TCUPS (ours / it) |")
    print("|---|---|---|---|---|---|---|---|")
    same_box = d.get("reference_cuda_same_box") or {}

    def row(name, workload, value, ms, e2e, frac, parity, against, kernel):
        ref = same_box.get(name, {})
        ref_text = f"{tcups(ref['gcups'])} ({ref.get('ours_over_reference_cuda')}x, equal={ref.get('matches_ours')})" if "gcups" in ref else "—"
        print(f"| {workload} | {kernel} | {tcups(value)} | {ms:.3f} | {tcups(e2e)} | {frac} | {parity} ({against}) | {ref_text} |")

    main_name = d["config"]["workload"].split(":")[0]
    row(main_name, d["config"]["workload"], d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("roofline") or {}).get("frac"),
        d["parity"], d["parity_against"], (d.get("roofline") or {}).get("kernel"))
    if d.get("dpx_tier"):
        t = d["dpx_tier"]
        print(f"| same tile pinned to the int16 DPX tier | szb::wavefront_kernel | {tcups(t['value'])} | {t['ms_per_step']:.3f} | — | {(t.get('roofline') or {}).get('frac')} | equal to the bit-parallel matrix | — |")
    if d.get("e2e_managed"):
        print(f"| same tile, managed-memory inputs and results | | — | {d['e2e_managed']['ms_per_step']:.3f} | {tcups(d['e2e_managed']['value'])} | | | |")
    for name, w in (d.get("workloads") or {}).items():
        row(name, w["workload"] + (f" [{w['scaling']}]" if n > 1 else ""), w["value"], w["ms_per_step"], w["e2e"]["value"], w.get("roofline_frac"),
            w["parity"], w["against"], w["kernel"])
    if d.get("cpu_baseline"):
        b = d["cpu_baseline"]
        print(f"\nreference CPU library on this box: {tcups(b['value'])} TCUPS, {b['cores']} threads ({b['sample']})")
