"""CUDA-graph capture of a whole training step body (forward + loss + backward [+ fused optimizer]).

At B200 speeds the step of the default model is a few dozen kernels lasting about a millisecond in total, so
Python/autograd/ctypes launch overhead becomes visible; capturing the step once and replaying it removes the host from
the loop (the reference's own loop syncs the host every step through ``l.item()``, main.py:170).

    step = GraphedStep(model, batch_shape=(B, 1, 12, N), target_shape=(B, N))
    loss = step(x, y)          # x, y: CUDA tensors (copied into the static buffers); gradients land in p.grad

Dropout: the kernels take their seed by value, which a capture would freeze; the step therefore owns a device-side
step counter, registered with the library (``stgcn_set_dropout_step``) and incremented by the last node of the graph,
so every replay draws fresh masks and the forward and backward of one replay agree.

Data parallelism: with ``reducer=FlatGradAllReducer(model)`` the gradients are views into one flat buffer that the
captured backward kernels write directly, and the ``ncclAvg`` all-reduce of that buffer is enqueued right behind every
replay (no pack/unpack copies, no scaling kernel).  Capturing the collective INSIDE the graph (bucket 0 issued from an
autograd hook behind ``st_blocks[1]``'s backward so that it overlaps ``st_blocks[0]``'s) was built and tried on 2 B200s
in round 2: both users of it (the 2-GPU test and bench.py) dead-locked at the first replays
(profiles/r02_ab_batch_d.md), so the collective stays outside the graph; measured 96 % weak-scaling efficiency at
2 GPUs with it there.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import torch

from . import _lib as L


class GraphedStep:
    def __init__(self, model: torch.nn.Module, batch_shape, target_shape, device=None,
                 post_backward: Optional[Callable[[], None]] = None, warmup: int = 3, reducer=None):
        """reducer: a dist.FlatGradAllReducer (None = single process); it runs after every replay.
        post_backward: called at the end of the captured body (e.g. a fused optimizer step, optim.FlatAdamW.step; with a
        reducer the optimizer must run after the all-reduce, i.e. outside: call it after ``step(x, y)``).
        (Splitting the batch into chains on parallel streams inside the graph was measured and removed: 2 chains -8 %,
        4 chains -30 % on PeMSD7-M B=256 -- the persistent kernels of the chains compete for the same SMs,
        profiles/r02_ab_batch_b.md.)"""
        self.model = model
        dev = device or next(model.parameters()).device
        self.device = dev
        self.x = torch.zeros(batch_shape, device=dev)
        self.y = torch.zeros(target_shape, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.post_backward = post_backward
        self.reducer = reducer
        self._lib = L.lib()
        self.step_counter = torch.zeros(1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            torch.cuda.synchronize(dev)
            L.check(self._lib.stgcn_set_dropout_step(self.step_counter.data_ptr()))
        # warm up on a side stream (allocator pools, lazily sized workspaces, cuFuncSetAttribute calls, helper streams)
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for i in range(max(warmup, 2 if reducer is not None else 1)):
                self._body()
                if reducer is not None:
                    reducer()                   # first call binds the flat buffer; later warm-ups run NCCL eagerly
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        for p in model.parameters():
            p.grad = None
        # thread_local: other threads (NCCL's watchdog, when a process group exists) may call into the runtime meanwhile
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self._body()
        self.params = list(model.parameters())
        self.grads = [p.grad for p in self.params]

    def _body(self):
        model = self.model
        for p in model.parameters():
            p.grad = None
        B = self.x.shape[0]
        pred = model(self.x).reshape(B, -1).float()
        dpred = torch.empty_like(pred)
        L.check(self._lib.stgcn_mse_fwd_bwd(pred.data_ptr(), self.y.data_ptr(), pred.numel(), C.c_float(1.0),
                                            self.loss.data_ptr(), dpred.data_ptr(),
                                            torch.cuda.current_stream(self.x.device).cuda_stream))
        pred.backward(dpred)
        if self.post_backward is not None:
            self.post_backward()
        self.step_counter.add_(1)

    def _after_replay(self):
        # a zero_grad(set_to_none=True) between steps detaches p.grad from the tensors the graph writes: re-point them
        for p, g in zip(self.params, self.grads):
            p.grad = g
        if self.reducer is not None:
            self.reducer()

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        self._after_replay()
        return self.loss

    def replay(self) -> torch.Tensor:
        """Replay on whatever is already in the static buffers ``self.x`` / ``self.y``."""
        self.graph.replay()
        self._after_replay()
        return self.loss

    def close(self) -> None:
        """Unregister the device step counter (the object must not be replayed afterwards)."""
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            L.check(self._lib.stgcn_set_dropout_step(None))
