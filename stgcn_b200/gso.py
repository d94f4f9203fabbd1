"""Graph-shift-operator preprocessing on the device: dense-tensor counterparts of the reference's
``calc_gso`` / ``calc_chebynet_gso`` (script/utility.py:6-76), same names and argument meaning, CUDA tensors in and out."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

GSO_TYPES = {"sym_norm_adj": 0, "sym_renorm_adj": 1, "sym_norm_lap": 2, "sym_renorm_lap": 3,
             "rw_norm_adj": 4, "rw_renorm_adj": 5, "rw_norm_lap": 6, "rw_renorm_lap": 7}


def _build(adj: torch.Tensor, code: int, chebynet: bool):
    if not adj.is_cuda or adj.dtype != torch.float32 or adj.dim() != 2 or adj.shape[0] != adj.shape[1]:
        raise RuntimeError("stgcn_b200.gso: expected a square float32 CUDA tensor")
    adj = adj.contiguous()
    n = adj.shape[0]
    out = torch.empty_like(adj)
    eig = torch.zeros(2, dtype=torch.float32, device=adj.device)
    ws = torch.empty(n * n + 3 * n + 8, dtype=torch.float32, device=adj.device)
    with torch.cuda.device(adj.device):
        L.check(L.lib().stgcn_gso_build(adj.data_ptr(), n, code, int(chebynet), out.data_ptr(), eig.data_ptr(),
                                        ws.data_ptr(), ws.numel(), torch.cuda.current_stream(adj.device).cuda_stream))
    return out, eig


def calc_gso(dir_adj: torch.Tensor, gso_type: str) -> torch.Tensor:
    """utility.py:6-57 on a dense (N, N) adjacency: symmetrise, (re)normalise, optionally form the Laplacian."""
    if gso_type not in GSO_TYPES:
        raise ValueError(f"{gso_type} is not defined.")                                   # utility.py:54
    return _build(dir_adj, GSO_TYPES[gso_type], False)[0]


def calc_chebynet_gso(gso: torch.Tensor, return_eigval: bool = False):
    """utility.py:59-76: 2 L / lambda_max - I (or L - I when lambda_max >= 2), lambda_max = ||L||_2 by power iteration."""
    if not gso.is_cuda or gso.dtype != torch.float32 or gso.dim() != 2 or gso.shape[0] != gso.shape[1]:
        raise RuntimeError("stgcn_b200.gso: expected a square float32 CUDA tensor")
    gso = gso.contiguous()
    n = gso.shape[0]
    lib = L.lib()
    # the rescale alone: feed the operator through the builder's last two stages by treating it as already normalised
    out = torch.empty_like(gso)
    eig = torch.zeros(2, dtype=torch.float32, device=gso.device)
    ws = torch.empty(n * n + 3 * n + 8, dtype=torch.float32, device=gso.device)
    with torch.cuda.device(gso.device):
        L.check(lib.stgcn_gso_rescale(gso.data_ptr(), n, out.data_ptr(), eig.data_ptr(), ws.data_ptr(), ws.numel(),
                                      torch.cuda.current_stream(gso.device).cuda_stream))
    return (out, eig) if return_eigval else out


def build_operator(dir_adj: torch.Tensor, gso_type: str, chebynet: bool) -> torch.Tensor:
    """calc_gso followed (for Chebyshev convolutions, main.py:97-101) by calc_chebynet_gso, in one call."""
    if gso_type not in GSO_TYPES:
        raise ValueError(f"{gso_type} is not defined.")
    return _build(dir_adj, GSO_TYPES[gso_type], chebynet)[0]
