#!/usr/bin/env python3
"""Small workload for compute-sanitizer: N=1 loop-back and 2 ranks on one device,
reach-only and a tiny sliced probe, both paths.  Parity is asserted so a sanitizer-clean run is also a
correct one."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cdprobe_pkg  # noqa: E402
from oracle import oracle as o  # noqa: E402

pkg = cdprobe_pkg.load()
SEED = o.DEFAULT_SEED
for path in (0, 1, 2):  # TMA bulk, 128-bit ld/st, 256-bit ld/st
    for n, mode, nbytes, extra in ((1, 1, (1 << 20) + 128 * 5, 0), (2, 0, 1 << 20, 0), (2, 1, (1 << 19) + 128 * 3, 0),
                                   (3, 1, 3 << 18, 0), (4, 1, 1 << 19, 0x80), (4, 1, 1 << 19, 0x20),
                                   (4, 1, 1 << 19, 0x400), (4, 1, 1 << 19, 0x800), (5, 1, 5 << 17, 0)):  # all-rank / pair barriers; odd domain
        flags = extra | (0x40 | 0x10 if n > 1 else 0)
        with pkg.Open(pkg.Config(ordinals=[0] * n, bytes=nbytes, mode=mode, flags=flags, ctas=4, timeout_ms=120000)) as p:
            p.SetOption(pkg.abi.OPT_PATH, path)
            p.SetOption(pkg.abi.OPT_VERIFY_CTAS, 1)
            p.SetOption(pkg.abi.OPT_WARMUP, 2)  # exercise the wake-up phase too
            if n >= 3:
                p.SetOption(pkg.abi.OPT_CTAS_RANK, (1 << 16) | 1)  # mixed grids: rank 0 on one CTA (both jobs on CTA 0)
            for _ in range(2):
                r = p.Run()
                assert not r.aborted
                for i in range(n):
                    for j in range(n):
                        if i == j and n > 1:
                            continue
                        assert r.reach_read[i][j] == 1 and r.reach_write[i][j] == 1, (n, mode, i, j)
                        assert (r.sum_read[i][j], r.xor_read[i][j]) == o.expected_read(SEED, n, nbytes, mode, i, j)
        print("ok", ("tma", "ldst", "ldst256")[path], n, mode, nbytes, hex(extra), flush=True)
print("SANITIZE_TARGET_DONE")
