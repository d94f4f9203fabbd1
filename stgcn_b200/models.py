"""Model assembly mirroring hazdzz/STGCN ``model/models.py`` (STGCNChebGraphConv :6-53, STGCNGraphConv :55-103).

The reference's own ``models.py`` loads ``stgcn_b200.layers`` unchanged; this module exists so
that bench.py / smoke() / the GPU tests can build the same network where ``/root/reference`` is not mounted.  The
attribute names (``st_blocks``, ``output``, ``fc1``, ``fc2``, ``Ko``) and therefore the state_dict keys are the
reference's.
"""
from __future__ import annotations

import torch.nn as nn

from . import layers


class _STGCNBase(nn.Module):
    """``blocks`` = [[c_in], [c1, c2, c3] * n_st, [o0, o1] (or [o0] when Ko == 0), [c_end]] (main.py:82-92);
    ``args`` carries Kt, Ks, act_func, graph_conv_type, gso, enable_bias, droprate, n_his (main.py:39-63,103)."""

    _dropout_attr = "dropout"

    def __init__(self, args, blocks, n_vertex):
        super().__init__()
        n_st = len(blocks) - 3
        self.st_blocks = nn.Sequential(*[
            layers.STConvBlock(args.Kt, args.Ks, n_vertex, blocks[l][-1], blocks[l + 1], args.act_func,
                               args.graph_conv_type, args.gso, args.enable_bias, args.droprate)
            for l in range(n_st)])
        self.Ko = args.n_his - n_st * 2 * (args.Kt - 1)            # models.py:34
        if self.Ko > 1:
            self.output = layers.OutputBlock(self.Ko, blocks[-3][-1], blocks[-2], blocks[-1][0], n_vertex,
                                             args.act_func, args.enable_bias, args.droprate)
        elif self.Ko == 0:                                          # models.py:38-42 (parameters only, see forward)
            self.fc1 = nn.Linear(in_features=blocks[-3][-1], out_features=blocks[-2][0], bias=args.enable_bias)
            self.fc2 = nn.Linear(in_features=blocks[-2][0], out_features=blocks[-1][0], bias=args.enable_bias)
            self.relu = nn.ReLU()
            setattr(self, self._dropout_attr, nn.Dropout(p=args.droprate))

    def forward(self, x):
        x = self.st_blocks(x)
        if self.Ko > 1:
            x = self.output(x)
        elif self.Ko == 0:
            # Unreachable in the reference as well: Ko == 0 means the last temporal conv sees fewer than Kt
            # steps and raises before this point (models.py:46-51); the ST blocks above raise the same way.
            raise RuntimeError("STGCN: Ko == 0 leaves no time steps for the output stage")
        return x                                                    # Ko == 1: raw ST-block output (models.py:46-51)


class STGCNChebGraphConv(_STGCNBase):
    """'TGTND TGTND TNFF' with Chebyshev graph convolutions (models.py:6-53)."""


class STGCNGraphConv(_STGCNBase):
    """'TGTND TGTND TNFF' with first-order (GCN) graph convolutions (models.py:55-103)."""
    _dropout_attr = "do"                                            # models.py:92
