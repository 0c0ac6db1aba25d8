#!/usr/bin/env python3
"""The workload `ncu` is pointed at for NVLink counters (VERDICT r01 missing #4).

ncu serialises and replays kernels, so a probe whose kernels wait for each other across GPUs cannot be
captured.  CDPROBE_OPT_SOLO_RANK runs ONE rank's own transfers in one self-contained kernel (no cross-GPU
barrier, no verify on the peer); with --ops 1 that kernel only reads its partner's slices, with --ops 2 it only
writes its partner's landing slots.  cdprobe_kernel launches of this process: 2 at open (local source
checksums, one per rank), then one per solo run — capture with `-k regex:cdprobe_kernel --launch-skip 3 --launch-count 1`.

    ncu --metrics nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,... \
        --clock-control none -k regex:cdprobe_kernel --launch-skip 3 --launch-count 1 --csv --log-file out.csv \
        python tools/solo_profile.py --ops 2
"""
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
ap.add_argument("--ops", type=int, default=3)
ap.add_argument("--runs", type=int, default=3)
ap.add_argument("--path", type=int, default=0)
a = ap.parse_args()
with pkg.Open(pkg.Config(ordinals=list(range(a.gpus)), bytes=a.bytes, ops=a.ops, timeout_ms=20000)) as p:
    p.SetOption(pkg.abi.OPT_WARMUP, 0)
    p.SetOption(pkg.abi.OPT_PATH, a.path)
    p.SetOption(pkg.abi.OPT_SOLO_RANK, 1)
    for _ in range(a.runs):
        r = p.Run()
    n = a.gpus
    print(json.dumps({"n": n, "ops": a.ops, "bytes_per_pair": r.bytes_per_pair, "payload_bytes_per_op": r.bytes_per_pair * (n - 1),
                      "read_gbps": [r.gbps_read[0][j] for j in range(1, n)], "write_gbps": [r.gbps_write[0][j] for j in range(1, n)],
                      "reach_read": [r.reach_read[0][j] for j in range(1, n)], "device_ms": r.device_ms[0]}))
