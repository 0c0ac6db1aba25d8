#!/usr/bin/env python3
"""BASELINE config 5 — reconcile storm: one handle, `--cycles` x { cdprobe_run(sliced);
emulate NodeUnprepare/NodePrepare by unmapping + remapping one peer }.  Reports p50/p99 per
cycle and checks for leaks (device memory via cudaMemGetInfo, process fds) — SURVEY H10 / T5."""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cdprobe_pkg  # noqa: E402


def n_fds():
    return len(os.listdir("/proc/self/fd"))


def main():
    import torch

    pkg = cdprobe_pkg.load()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--cycles", type=int, default=1000)
    ap.add_argument("--bytes", type=int, default=1 << 30)
    ap.add_argument("--same-device", action="store_true", help="put all ranks on GPU 0 (1-GPU box)")
    args = ap.parse_args()
    n = args.gpus
    flags = (0x40 | 0x10) if args.same_device else 0
    ords = [0] * n if args.same_device else list(range(n))
    with pkg.Open(pkg.Config(ordinals=ords, bytes=args.bytes, flags=flags, timeout_ms=20000,
                             ctas=8 if args.same_device else 0)) as p:
        for _ in range(5):
            p.Run()
        free0 = [torch.cuda.mem_get_info(d)[0] for d in sorted(set(ords))]
        fd0 = n_fds()
        run_ms, cyc_ms, bad = [], [], 0
        t_all = time.perf_counter()
        for c in range(args.cycles):
            t0 = time.perf_counter()
            r = p.Run()
            run_ms.append(r.probe_ms)
            if not all(all(x == 1 for x in row) for row in r.reach):
                bad += 1
            if n > 1:
                a = c % n
                b = (a + 1 + (c // n) % (n - 1)) % n
                p.RemapPeer(a, b)  # unmap + map again: the "unprepare/prepare" of one peer
            cyc_ms.append((time.perf_counter() - t0) * 1e3)
        wall = time.perf_counter() - t_all
        free1 = [torch.cuda.mem_get_info(d)[0] for d in sorted(set(ords))]
        fd1 = n_fds()
    q = lambda v, f: sorted(v)[min(len(v) - 1, int(f * len(v)))]
    print(json.dumps({
        "config": "reconcile storm", "n_gpus": n, "cycles": args.cycles, "bytes": args.bytes,
        "probe_ms_p50": statistics.median(run_ms), "probe_ms_p99": q(run_ms, 0.99), "probe_ms_max": max(run_ms),
        "cycle_ms_p50": statistics.median(cyc_ms), "cycle_ms_p99": q(cyc_ms, 0.99), "wall_s": wall,
        "unreachable_runs": bad, "device_free_delta_bytes": [a - b for a, b in zip(free0, free1)],
        "fd_delta": fd1 - fd0}))
    return 0 if bad == 0 and fd1 == fd0 and all(a == b for a, b in zip(free0, free1)) else 1


if __name__ == "__main__":
    sys.exit(main())
