"""world_size-2 gloo test (CPU) of the batch sharding + flat gradient all-reduce: averaged shard gradients of
the oracle model equal the full-batch gradients (MSE is a mean, LayerNorm is per sample)."""
import os
import sys
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _OracleNet(torch.nn.Module):
    """The oracle wrapped as an nn.Module (test infrastructure; the product layers are CUDA-only)."""

    def __init__(self, params, gso, cfg):
        super().__init__()
        self.keys = list(params)
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(v.clone()) for v in params.values()])
        self.gso, self.cfg = gso, cfg

    def forward(self, x):
        from oracle import stgcn_oracle as O
        return O.stgcn_forward(x, dict(zip(self.keys, self.ps)), self.gso, **self.cfg)


def _worker(rank, world, init_file, out_file):
    sys.path.insert(0, ROOT)
    from oracle import stgcn_oracle as O
    from stgcn_b200.dist import FlatGradAllReducer, shard_batch
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    torch.manual_seed(0)
    n, B = 15, 8
    blocks = [[1], [8, 4, 8], [8, 4, 8], [16, 16], [1]]
    cfg = dict(blocks=blocks, kt=3, n_his=12, act="glu", kind="cheb_graph_conv")
    gso = O.synthetic_gso(n, seed=1)
    params = O.init_params(blocks=blocks, kt=3, ks=3, n_his=12, n_vertex=n, seed=3)
    x = torch.randn(B, 1, 12, n)
    y = torch.randn(B, n)
    net = _OracleNet(params, gso, cfg)
    red = FlatGradAllReducer(net)
    sl = shard_batch(B, rank, world)
    for _ in range(2):   # two steps: the second exercises the cached flat buffer
        net.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(net(x[sl]).reshape(sl.stop - sl.start, -1), y[sl])
        loss.backward()
        red()
    if rank == 0:
        full = _OracleNet(params, gso, cfg)
        torch.nn.functional.mse_loss(full(x).reshape(B, -1), y).backward()
        worst = 0.0
        n_live = 0
        for a, b in zip(net.ps, full.ps):
            assert (a.grad is None) == (b.grad is None)
            if b.grad is not None:
                n_live += 1
                worst = max(worst, float((a.grad - b.grad).norm() / b.grad.norm().clamp_min(1e-30)))
        torch.save({"worst": worst, "n_live": n_live, "n_all": len(net.ps), "flat": red.numel}, out_file)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_allreduce_matches_full_batch():
    with tempfile.TemporaryDirectory() as td:
        init_file, out_file = os.path.join(td, "init"), os.path.join(td, "out.pt")
        mp.spawn(_worker, args=(2, init_file, out_file), nprocs=2, join=True)
        res = torch.load(out_file)
    assert res["worst"] < 1e-5
    assert res["n_live"] == res["n_all"] - 10          # the 10 dead align-conv tensors are skipped
    assert res["flat"] > 0


def test_shard_batch():
    from stgcn_b200.dist import shard_batch
    import pytest
    assert [shard_batch(8, r, 4) for r in range(4)] == [slice(0, 2), slice(2, 4), slice(4, 6), slice(6, 8)]
    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)
