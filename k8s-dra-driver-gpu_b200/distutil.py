"""Rank-group helpers for the one-process-per-GPU layout (bench.py under torchrun).

The probe's data path never touches torch.distributed (NVLink P2P + device flag barrier);
this only brackets timed regions and reduces a handful of host scalars.  Backend "nccl" on
the GPU box, "gloo" in the CPU tests (world_size 2).
"""
from __future__ import annotations

import os
from typing import Optional


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


class RankGroup:
    def __init__(self, backend: Optional[str] = None, device=None):
        self.rank, self.world, self.local = dist_env()
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch.distributed as dist

            kw = {}
            if backend == "nccl" and device is not None:
                kw["device_id"] = device
            dist.init_process_group(backend=backend or "gloo", **kw)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, x: float, op):
        import torch

        t = torch.tensor([x], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.dist is not None else float(x)

    def min(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MIN) if self.dist is not None else float(x)

    def gather_objects(self, obj):
        """Every rank's small picklable object, in rank order, on every rank."""
        if self.dist is None:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def session(self, prefix: str = "bench") -> str:
        """A rendezvous name every rank of this launch derives identically."""
        return f"{prefix}-{os.environ.get('MASTER_PORT', '0')}-{os.environ.get('TORCHELASTIC_RUN_ID', 'x')}"

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()
            self.dist = None
