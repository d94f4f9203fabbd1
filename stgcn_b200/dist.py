"""Batch-sharded data parallelism for the ST-block path: one process per GPU, replicated parameters and operator,
contiguous batch shards, and ONE flat fp32 gradient buffer that the backward kernels write into directly and that is
averaged across ranks in place.  The reference has no distributed code at all; vanilla
DistributedDataParallel also fails on it because 10 of 38 parameters (the dead align convs, layers.py:12) never
receive a gradient.  The reducer therefore works on the parameters that actually produce a gradient.

Backend-agnostic (NCCL on the B200s over NVLink/NVSwitch, gloo in the CPU tests).

Layout of the flat buffer: the parameters are ordered by the time their gradient becomes final in the backward pass --
output stage first, then the ST blocks from the last to the first -- and cut into ``buckets``: bucket 0 = everything but
the first ST block, bucket 1 = the first ST block.  ``reduce_bucket(0)`` can therefore be issued as soon as the backward
of ``st_blocks[1]`` has been enqueued and overlaps the backward of ``st_blocks[0]``; the data-path collective is still
one all-reduce of every live gradient per step, split at one point."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_batch(n_samples: int, rank: int, world: int) -> slice:
    """Contiguous equal shards of the global batch (rank r gets [r*B/W, (r+1)*B/W))."""
    if n_samples % world:
        raise ValueError(f"global batch {n_samples} is not divisible by world size {world}")
    per = n_samples // world
    return slice(rank * per, (rank + 1) * per)


def _backward_order(module: torch.nn.Module) -> List[Tuple[str, torch.nn.Parameter]]:
    """Named parameters in the order their gradients become final: everything outside ``st_blocks`` (the output stage)
    first, then ``st_blocks.<l>`` for l descending."""
    named = list(module.named_parameters())

    def key(item):
        name = item[0]
        if name.startswith("st_blocks."):
            return (1, -int(name.split(".")[1]))
        return (0, 0)

    return sorted(named, key=key)          # stable: registration order inside a group


class FlatGradAllReducer:
    """Gradients of the live parameters as views into one flat fp32 buffer, averaged with (at most two) all-reduces.

    Usage::

        reducer = FlatGradAllReducer(model)
        loss.backward()                 # first step: ordinary gradients; discovers which parameters are live
        reducer()                       # binds the flat buffer (copying this step's gradients in) and reduces
        ...
        model.zero_grad(set_to_none=True); loss.backward(); reducer()      # later steps: the stgcn_b200 backward
                                        # kernels write straight into the flat buffer (layers._grad_like), no copies

    The set of live parameters is a property of the architecture, hence identical on every rank.  ``p.grad`` must be
    ``None`` (``zero_grad(set_to_none=True)``) before every backward once the buffer is bound: autograd then adopts the
    view the kernels filled; with a stale ``p.grad`` it would add the new gradient to itself."""

    ALIGN = 64          # floats

    def __init__(self, module: torch.nn.Module, group: Optional[dist.ProcessGroup] = None, split_first_block: bool = True):
        self.module = module
        self.group = group
        self.split_first_block = split_first_block
        self.live: Optional[List[torch.nn.Parameter]] = None
        self.names: List[str] = []
        self.flat: Optional[torch.Tensor] = None
        self.sizes: List[int] = []
        self.offsets: List[int] = []
        self.bucket_bounds: List[Tuple[int, int]] = []       # element ranges of the buckets inside ``flat``
        self._pending: Dict[int, object] = {}

    # ------------------------------------------------------------------ setup
    def bind(self) -> None:
        """Discover the live parameters (those holding a gradient now), allocate the flat buffer, copy the current
        gradients in and re-point ``p.grad`` at the views."""
        order = [(n, p) for n, p in _backward_order(self.module) if p.grad is not None]
        if not order:
            raise RuntimeError("FlatGradAllReducer.bind(): no parameter has a gradient yet (run one backward first)")
        self.names = [n for n, _ in order]
        self.live = [p for _, p in order]
        self.sizes = [p.numel() for p in self.live]
        # every slot starts on a 256-byte boundary (64 floats), like a tensor of its own would: the kernels take their
        # 16-byte vector paths only on aligned parameter / gradient pointers.  The padding stays zero.
        self.offsets, total = [], 0
        for s in self.sizes:
            self.offsets.append(total)
            total += (s + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        dev = self.live[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        cut = total
        if self.split_first_block:
            first = [i for i, n in enumerate(self.names) if n.startswith("st_blocks.0.")]
            if first and first[0] > 0:
                cut = self.offsets[first[0]]
        self.bucket_bounds = [(0, cut)] + ([(cut, total)] if cut < total else [])
        for p, off, n in zip(self.live, self.offsets, self.sizes):
            view = self.flat.narrow(0, off, n).view_as(p)
            view.copy_(p.grad)
            p.grad = view
            p._stgcn_grad_slot = (self.flat, off, n)       # layers._grad_like hands out fresh views of this slot

    def unbind(self) -> None:
        for p in self.live or []:
            if hasattr(p, "_stgcn_grad_slot"):
                del p._stgcn_grad_slot
        self.live = None

    @property
    def numel(self) -> int:
        return 0 if self.flat is None else self.flat.numel()

    @property
    def n_buckets(self) -> int:
        return len(self.bucket_bounds)

    def _world(self) -> int:
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.group)

    # ------------------------------------------------------------------ reduction
    def reduce_bucket(self, b: int, async_op: bool = False) -> None:
        """Average bucket ``b`` of the flat buffer across ranks, in place.  With ``async_op`` the collective is only
        enqueued (on the backend's own stream, ordered after the work enqueued so far on the current stream); ``wait()``
        joins it back."""
        world = self._world()
        if world == 1 or self.flat is None:
            return
        lo, hi = self.bucket_bounds[b]
        chunk = self.flat.narrow(0, lo, hi - lo)
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            work = dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        else:                                   # gloo has no AVG
            work = dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
            if not async_op:
                chunk.mul_(1.0 / world)
        if async_op:
            self._pending[b] = (work, chunk, backend != "nccl", world)

    def wait(self) -> None:
        for b in sorted(self._pending):
            work, chunk, scale, world = self._pending[b]
            work.wait()
            if scale:
                chunk.mul_(1.0 / world)
        self._pending.clear()

    def __call__(self) -> None:
        """Reduce everything now (binds the flat buffer on first use)."""
        if self._world() == 1:
            return
        if self.live is None:
            self.bind()
        else:
            self._check_bound()
        for b in range(self.n_buckets):
            if b not in self._pending:
                self.reduce_bucket(b)
        self.wait()

    def _check_bound(self) -> None:
        """Gradients that do not live in the flat buffer (a backward that did not go through stgcn_b200's kernels, or a
        stale ``p.grad``) are copied in -- correct, merely slower."""
        for p, off, n in zip(self.live, self.offsets, self.sizes):
            g = p.grad
            if g is None:
                raise RuntimeError("a parameter that had a gradient on the first step has none now")
            if g.data_ptr() != self.flat.data_ptr() + 4 * off:
                view = self.flat.narrow(0, off, n).view_as(p)
                view.copy_(g)
                p.grad = view
