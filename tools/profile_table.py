#!/usr/bin/env python3
"""Prints the round summary table from profiles/<prefix>_final_bench_n*.json (so docs never drift from files)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prefix = sys.argv[1] if len(sys.argv) > 1 else "r01"


def load(name):
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        return None
    return json.loads([l for l in open(p) if l.startswith("{")][-1])


print("| N | probe ms (CUDA events) | e2e ms (host call) | pair GB/s bidirectional read · write | one-way read · write (probe ms) | cold probe after 1 s idle | reference CPU poll (mean / median) |")
print("|---|---|---|---|---|---|---|")
for n in (1, 2, 4, 8):
    b, r = load(f"{prefix}_final_bench_n{n}.json"), load(f"{prefix}_final_ref_n{n}.json")
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
    ref = "—" if not r else f"{r['value']:.0f} / {r['cpu_baseline']['median_ms']:.0f} ms"
    print(f"| {n} | {probe} | {b['e2e']['value']:.3f} | {pair} | {uni} | {b['cold_start']['probe_ms']:.2f} | {ref} |")
