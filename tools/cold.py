#!/usr/bin/env python3
"""Cold-start experiment: how slow is the first NVLink phase after an idle gap, and how many
wake-up bytes cure it?  Prints one JSON line per (gap, warm bytes)."""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cdprobe_pkg  # noqa: E402

pkg = cdprobe_pkg.load()
abi = pkg.abi
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=2)
ap.add_argument("--bytes", type=int, default=1 << 30)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cold.jsonl"))
a = ap.parse_args()
n = a.gpus
with pkg.Open(pkg.Config(ordinals=list(range(n)), bytes=a.bytes, timeout_ms=20000)) as p, open(a.out, "w") as f:
    for _ in range(3):
        p.Run()
    for gap in (0.0, 0.0005, 0.002, 0.01, 0.05, 0.2, 1.0, 3.0):
        for warm in (0, 8 << 20, 32 << 20, 64 << 20, 128 << 20, 256 << 20):
            p.SetOption(abi.OPT_WARMUP, 2 if warm else 0)
            p.SetOption(abi.OPT_WARMUP_BYTES, warm)
            rec = []
            for rep in range(3):
                time.sleep(gap)
                r = p.Run()
                tr = p.Trace(0)
                first = next(ph for ph in tr if ph["job0"] == "read")
                d = first["t_end0"] - first["t_start"]
                off = [(i, j) for i in range(n) for j in range(n) if i != j]
                rec.append({"first_read_gbps": r.bytes_per_pair / d, "min_read": min(r.gbps_read[i][j] for i, j in off),
                            "min_write": min(r.gbps_write[i][j] for i, j in off), "probe_ms": r.probe_ms,
                            "warm_us": (tr[0]["t_arrive"] - tr[0]["t_start"]) / 1e3})
            out = {"n": n, "gap_s": gap, "warm_mib": warm >> 20,
                   **{k: statistics.median(x[k] for x in rec) for k in rec[0]},
                   "min_read_worst": min(x["min_read"] for x in rec)}
            f.write(json.dumps(out) + "\n")
            print(json.dumps(out), flush=True)
    # default behaviour (auto) after a long gap
    p.SetOption(abi.OPT_WARMUP, 1)
    p.SetOption(abi.OPT_WARMUP_BYTES, 128 << 20)
    time.sleep(1.0)
    r = p.Run()
    print("auto after 1 s idle: warmed", r.warmed, "min read", min(r.gbps_read[i][j] for i in range(n) for j in range(n) if i != j), "probe_ms", r.probe_ms)
    r = p.Run()
    print("auto back-to-back: warmed", r.warmed, "probe_ms", r.probe_ms)
