"""CUDA-graph capture of a whole training step body (forward + loss + backward [+ gradient all-reduce]).

At B200 speeds the step of the default model is a few hundred small kernels lasting a few milliseconds in total, so
Python/autograd/ctypes launch overhead becomes visible; capturing the step once and replaying it removes the host from
the loop (the reference's own loop syncs the host every step through ``l.item()``, main.py:170).

    step = GraphedStep(model, batch_shape=(B, 1, 12, N), target_shape=(B, N))
    loss = step(x, y)          # x, y: CUDA tensors (copied into the static buffers); gradients land in p.grad
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import torch

from . import _lib as L


class GraphedStep:
    def __init__(self, model: torch.nn.Module, batch_shape, target_shape, device=None,
                 post_backward: Optional[Callable[[], None]] = None, warmup: int = 3, micro_streams: int = 1):
        """micro_streams > 1 (experimental, unmeasured at the end of round 1): the batch is cut into that many equal
        chunks whose forward+backward chains run on separate streams inside the ONE captured graph, so the tails and
        launch bubbles of one chain overlap the other's kernels; the chunk losses are scaled by 1/micro_streams, so
        the accumulated gradients and ``loss`` equal the full-batch ones.  Set STGCN_SIDE_PER_STREAM=1 as well (one
        pair of library helper streams per chain)."""
        self.model = model
        dev = device or next(model.parameters()).device
        self.x = torch.zeros(batch_shape, device=dev)
        self.y = torch.zeros(target_shape, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.post_backward = post_backward
        self.micro = int(micro_streams)
        if self.micro < 1 or batch_shape[0] % self.micro:
            raise ValueError(f"micro_streams={micro_streams} must divide the batch size {batch_shape[0]}")
        self._mstreams = [torch.cuda.Stream(device=dev) for _ in range(self.micro)] if self.micro > 1 else []
        self._mloss = [torch.zeros(1, device=dev) for _ in range(self.micro)] if self.micro > 1 else []
        self._lib = L.lib()
        # warm up on a side stream (allocator pools, lazily sized workspaces, cuFuncSetAttribute calls)
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        for p in model.parameters():
            p.grad = None
        with torch.cuda.graph(self.graph):
            self._body()
        self.grads = [p.grad for p in model.parameters()]

    def _body_micro(self):
        model = self.model
        for p in model.parameters():
            p.grad = None
        cur = torch.cuda.current_stream(self.x.device)
        k = self.micro
        for st, xc, yc, lc in zip(self._mstreams, self.x.chunk(k), self.y.chunk(k), self._mloss):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                Bc = xc.shape[0]
                pred = model(xc).reshape(Bc, -1).float()
                dpred = torch.empty_like(pred)
                L.check(self._lib.stgcn_mse_fwd_bwd(pred.data_ptr(), yc.data_ptr(), pred.numel(), C.c_float(1.0 / k),
                                                    lc.data_ptr(), dpred.data_ptr(), st.cuda_stream))
                pred.backward(dpred)
        for st in self._mstreams:
            cur.wait_stream(st)
        torch.stack(self._mloss).sum(0, out=self.loss)
        self.loss.mul_(1.0 / k)
        if self.post_backward is not None:
            self.post_backward()

    def _body(self):
        if self.micro > 1:
            return self._body_micro()
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

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.loss

    def replay(self) -> torch.Tensor:
        """Replay on whatever is already in the static buffers ``self.x`` / ``self.y``."""
        self.graph.replay()
        return self.loss
