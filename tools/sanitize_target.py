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
for path in (0, pkg.abi.FLAG_PATH_LDST):
    for n, mode, nbytes in ((1, 1, (1 << 20) + 128 * 5), (2, 0, 1 << 20), (2, 1, (1 << 19) + 128 * 3), (3, 1, 3 << 18)):
        flags = path | (0x40 | 0x10 if n > 1 else 0)
        with pkg.Open(pkg.Config(ordinals=[0] * n, bytes=nbytes, mode=mode, flags=flags, ctas=4, timeout_ms=120000)) as p:
            for _ in range(2):
                r = p.Run()
                assert not r.aborted
                for i in range(n):
                    for j in range(n):
                        if i == j and n > 1:
                            continue
                        assert r.reach_read[i][j] == 1 and r.reach_write[i][j] == 1, (n, mode, i, j)
                        assert (r.sum_read[i][j], r.xor_read[i][j]) == o.expected_read(SEED, n, nbytes, mode, i, j)
        print("ok", "ldst" if path else "tma", n, mode, nbytes, flush=True)
print("SANITIZE_TARGET_DONE")
