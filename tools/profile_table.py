#!/usr/bin/env python3
"""Prints the round summary table from the recorded bench lines (so docs never drift from files).

    python tools/profile_table.py r02      # profiles/r02_bench_n{1,2,4,8}.json + r02_ref_n*.json
    python tools/profile_table.py r01      # profiles/r01_final_bench_n*.json (round-1 naming)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prefix = sys.argv[1] if len(sys.argv) > 1 else "r02"
mid = "_final" if prefix == "r01" else ""


def load(name):
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        return None
    return json.loads([l for l in open(p) if l.startswith("{")][-1])


print("| N | probe ms (CUDA events) | e2e ms (host call) | flag barriers per probe | pair GB/s, both directions loaded: read · write | one way: read · write (probe ms) | same-box copy engine: one way · both ways | cold probe after 1 s idle | first verdict: open + cold run · fresh daemon process over the N GPUs | reference CPU poll, median (mean) |")
print("|---|---|---|---|---|---|---|---|---|---|")
for n in (1, 2, 4, 8):
    b, r = load(f"{prefix}{mid}_bench_n{n}.json"), load(f"{prefix}{mid}_ref_n{n}.json")
    if b is None:
        continue
    pl, u = b["per_link_gbps"], b.get("per_link_gbps_unidirectional")
    rf = b["roofline"]
    probe = f"{b['value']:.3f}"
    if n == 1:
        probe += f" (3 GiB HBM: {rf['achieved']:.0f} GB/s = {rf['frac']:.3f} of measured {rf['peak']:.0f})"
        pair = f"loop-back: write {pl['write_min']:.0f}, then read {pl['read_min']:.0f} concurrent with the verify"
    else:
        pair = f"{pl['read_min']:.0f}–{pl['read_max']:.0f} · {pl['write_min']:.0f}–{pl['write_max']:.0f}"
    uni = "—" if not u else f"{u['read_min']:.0f}–{u['read_median']:.0f} (min–median) · {u['write_min']:.0f}–{u['write_median']:.0f} ({u['probe_ms']:.2f})"
    ce = "—"
    if "ce" in rf:
        ce = f"{max(rf['ce']['uni_push'], rf['ce']['uni_pull']):.0f} (push {rf['ce']['uni_push']:.0f}) · {rf['ce']['bidi_push_min']:.0f}"
    bar = f"{b['barrier_us']:.0f} µs" if "barrier_us" in b else "—"
    dc = b.get("daemon_cost")
    first = f"{dc['cold_first_verdict_ms']:.0f} ms" if dc else "—"
    if dc and dc.get("daemon_process", {}).get("wall_ms"):
        first += f" · {dc['daemon_process']['wall_ms'] / 1e3:.1f} s"
    if not r:
        ref = "—"
    elif r["cpu_baseline"].get("statistic") == "median":
        ref = f"{r['value']:.0f} ({r['cpu_baseline']['mean_ms']:.0f}) ms"
    else:
        ref = f"{r['cpu_baseline']['median_ms']:.0f} ({r['value']:.0f}) ms"
    print(f"| {n} | {probe} | {b['e2e']['value']:.3f} | {bar} | {pair} | {uni} | {ce} | {b['cold_start']['probe_ms']:.2f} | {first} | {ref} |")
