#!/usr/bin/env python3
"""Dumps the per-phase device timeline (cdprobe_trace) of every local rank for one probe run."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cdprobe_pkg  # noqa: E402

pkg = cdprobe_pkg.load()
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=2)
ap.add_argument("--bytes", type=int, default=1 << 30)
ap.add_argument("--mode", default="sliced")
ap.add_argument("--flags", type=lambda x: int(x, 0), default=0)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trace.json"))
ap.add_argument("--idle", type=float, default=0.0, help="sleep this long before the traced run (cold start: the daemon's case)")
a = ap.parse_args()
mode = {"sliced": 1, "full": 2, "reach": 0}[a.mode]
with pkg.Open(pkg.Config(ordinals=list(range(a.gpus)), bytes=a.bytes, mode=mode, flags=a.flags, timeout_ms=20000)) as p:
    for _ in range(3):
        r = p.Run()
    if a.idle > 0:
        import time

        time.sleep(a.idle)
        r = p.Run()
    out = {"n": a.gpus, "mode": a.mode, "bytes": a.bytes, "flags": a.flags, "probe_ms": r.probe_ms,
           "device_ms": r.device_ms, "bpp": r.bytes_per_pair, "ranks": [p.Trace(i) for i in range(a.gpus)]}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"))
    for i, tr in enumerate(out["ranks"][:2]):
        print(f"rank {i}: probe {r.probe_ms:.3f} ms")
        for ph in tr:
            d0 = (ph["t_end0"] - ph["t_start"]) / 1e3 if ph["t_end0"] else 0
            d1 = (ph["t_end1"] - ph["t_start"]) / 1e3 if ph["t_end1"] else 0
            print(f"  {ph['job0']:6s}->{ph['peer0']:2d} {d0:8.1f} us | {ph['job1']:6s} {d1:8.1f} us | start {ph['t_start']/1e3:9.1f} arrive {ph['t_arrive']/1e3:9.1f} sync={ph['sync_mask']:#04x} post={ph['post_mask']:#04x}")
