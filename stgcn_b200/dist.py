"""Batch-sharded data parallelism for the ST-block path: one process per GPU, one flat gradient all-reduce
per step.  The reference has no distributed code at all; vanilla DistributedDataParallel
also fails on it because 10 of 38 parameters (the dead align convs, layers.py:12) never receive a gradient.
This reducer therefore works on the parameters that actually produced a gradient.

Backend-agnostic (NCCL on the B200s over NVLink/NVSwitch, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def shard_batch(n_samples: int, rank: int, world: int) -> slice:
    """Contiguous equal shards of the global batch (rank r gets [r*B/W, (r+1)*B/W))."""
    if n_samples % world:
        raise ValueError(f"global batch {n_samples} is not divisible by world size {world}")
    per = n_samples // world
    return slice(rank * per, (rank + 1) * per)


class FlatGradAllReducer:
    """Averages gradients across ranks with ONE all-reduce on a flat fp32 buffer.

    The set of live parameters (those with a gradient after the first backward) is fixed at the first call
    and must be identical on every rank -- it is, because it is a property of the architecture."""

    def __init__(self, module: torch.nn.Module, group: Optional[dist.ProcessGroup] = None):
        self.module = module
        self.group = group
        self.live: Optional[List[torch.nn.Parameter]] = None
        self.flat: Optional[torch.Tensor] = None
        self.sizes: List[int] = []

    def _setup(self):
        self.live = [p for p in self.module.parameters() if p.grad is not None]
        self.sizes = [p.numel() for p in self.live]
        dev = self.live[0].device
        self.flat = torch.empty(sum(self.sizes), dtype=torch.float32, device=dev)

    @property
    def numel(self) -> int:
        return 0 if self.flat is None else self.flat.numel()

    def __call__(self) -> None:
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(self.group)
        if world == 1:
            return
        if self.live is None:
            self._setup()
        grads = [p.grad for p in self.live]
        if any(g is None for g in grads):
            raise RuntimeError("a parameter that had a gradient on the first step has none now")
        views = list(self.flat.split(self.sizes))
        torch._foreach_copy_(views, [g.reshape(-1) for g in grads])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / world)
        torch._foreach_copy_([g.view(-1) for g in grads], views)
