"""Window construction on the device.  The reference's ``data_transform``
(script/dataloader.py:32-48) materialises every window of a split on the host -- ``num x 1 x n_his x N`` floats, 12x the
series -- and ships it to the GPU; here the z-scored series stays resident in HBM once and the windows of ONE batch are
gathered from it per step by a copy kernel (pure index work: bit exact)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L


class DeviceWindows:
    def __init__(self, series: torch.Tensor, n_his: int, n_pred: int):
        """series: (len, N) float32 CUDA tensor (already standardised, main.py:116-119)."""
        if not series.is_cuda or series.dtype != torch.float32 or series.dim() != 2:
            raise RuntimeError("DeviceWindows: expected a 2-D float32 CUDA tensor (len, N)")
        self.series = series.contiguous()
        self.n_his, self.n_pred = int(n_his), int(n_pred)
        self.len, self.N = self.series.shape

    def __len__(self) -> int:
        """Number of windows, exactly the reference's ``num`` (dataloader.py:37)."""
        return max(0, self.len - self.n_his - self.n_pred)

    def batch(self, start: int = 0, size: Optional[int] = None, starts: Optional[torch.Tensor] = None,
              out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(x, y) of the windows ``start .. start+size-1`` (the reference's sequential, unshuffled DataLoader order,
        main.py:126) or of the given window indices ``starts`` (int64 CUDA tensor).  x: (B, 1, n_his, N), y: (B, N)."""
        dev = self.series.device
        if starts is not None:
            starts = starts.to(device=dev, dtype=torch.int64).contiguous()
            B = starts.numel()
        else:
            B = min(size if size is not None else len(self), len(self) - start)
        if out is None:
            x = torch.empty((B, 1, self.n_his, self.N), dtype=torch.float32, device=dev)
            y = torch.empty((B, self.N), dtype=torch.float32, device=dev)
        else:
            x, y = out
        with torch.cuda.device(dev):
            L.check(L.lib().stgcn_windows(self.series.data_ptr(), self.len, self.N, self.n_his, self.n_pred,
                                          None if starts is None else starts.data_ptr(), int(start), B,
                                          x.data_ptr(), y.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        return x, y


class HostBatchPrefetcher:
    """Double-buffered host -> device staging of (x, y) batches on a copy stream, the device-side half of a pinned-memory
    data loader (the reference's DataLoader hands out host tensors that main.py:165 moves with ``.to(device)``: a blocking
    copy in front of every step).  ``request(i, x_host, y_host)`` starts the copy of batch ``i`` into staging pair ``i & 1``;
    ``take(i)`` makes the current stream wait for it and returns the device tensors; ``release(i)`` -- after the consumer's
    last read of them has been ENQUEUED on the current stream -- lets the copy of batch ``i + 2`` reuse the pair.  Typical
    step: ``x, y = pf.take(i); pf.request(i + 1, ...); step(x, y); pf.release(i)`` -- batch i+1 travels while step i computes.
    """

    def __init__(self, x_like: torch.Tensor, y_like: torch.Tensor, device):
        self.dev = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.stage = [(torch.empty(x_like.shape, dtype=x_like.dtype, device=self.dev),
                       torch.empty(y_like.shape, dtype=y_like.dtype, device=self.dev)) for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]
        self.free_valid = [False, False]
        self.requested = -1                       # index of the newest batch requested

    def request(self, i: int, x_host: torch.Tensor, y_host: torch.Tensor) -> None:
        if not (x_host.is_pinned() and y_host.is_pinned()):
            raise RuntimeError("HostBatchPrefetcher: host batches must be pinned (torch.Tensor.pin_memory())")
        k = i & 1
        with torch.cuda.stream(self.copy_stream):
            if self.free_valid[k]:
                self.copy_stream.wait_event(self.free[k])       # the previous user of this pair has consumed it
            self.stage[k][0].copy_(x_host, non_blocking=True)
            self.stage[k][1].copy_(y_host, non_blocking=True)
            self.ready[k].record(self.copy_stream)
        self.requested = i

    def take(self, i: int):
        if self.requested < i:
            raise RuntimeError("HostBatchPrefetcher.take: batch was never requested")
        k = i & 1
        torch.cuda.current_stream(self.dev).wait_event(self.ready[k])
        return self.stage[k]

    def release(self, i: int) -> None:
        k = i & 1
        self.free[k].record(torch.cuda.current_stream(self.dev))
        self.free_valid[k] = True
