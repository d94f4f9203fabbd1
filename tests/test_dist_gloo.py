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
    for it in range(3):   # step 0 binds the flat buffer; step 1 reduces it in one go; step 2 bucket by bucket, bucket 0 early
        net.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(net(x[sl]).reshape(sl.stop - sl.start, -1), y[sl])
        loss.backward()
        if it == 2:
            red._check_bound()
            red.reduce_bucket(0, async_op=True)         # what GraphedStep issues after st_blocks[1]'s backward
        red()
    assert all(p.grad.data_ptr() == red.flat.data_ptr() + 4 * off for p, off in zip(red.live, red.offsets))
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
        torch.save({"worst": worst, "n_live": n_live, "n_all": len(net.ps), "flat": red.numel,
                    "buckets": red.bucket_bounds, "names": red.names}, out_file)
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
    # the oracle wrapper's parameters are an unnamed list (no st_blocks.* names): one bucket; the bucket split itself is
    # covered by test_backward_order_and_buckets below
    assert res["buckets"] == [(0, res["flat"])]


def test_backward_order_and_buckets():
    """Flat-buffer layout on the real module tree (CPU: construction and binding only, no kernels): output stage first,
    then st_blocks.1, then st_blocks.0 in its own bucket; dead align convs are left out."""
    sys.path.insert(0, ROOT)
    from types import SimpleNamespace
    from stgcn_b200 import models
    from stgcn_b200.dist import FlatGradAllReducer
    n = 9
    blocks = [[1], [8, 4, 8], [8, 4, 8], [16, 16], [1]]
    args = SimpleNamespace(Kt=3, Ks=3, act_func="glu", graph_conv_type="cheb_graph_conv", gso=torch.eye(n),
                           enable_bias=True, droprate=0.0, n_his=12)
    model = models.STGCNChebGraphConv(args, blocks, n)
    dead = {k for k, _ in model.named_parameters() if ".align.align_conv." in k and "graph_conv" not in k}
    for k, p in model.named_parameters():
        if k not in dead:
            p.grad = torch.full_like(p, 2.0)
    red = FlatGradAllReducer(model)
    red.bind()
    assert red.names[0].startswith("output.")
    first = [i for i, k in enumerate(red.names) if k.startswith("st_blocks.0.")]
    second = [i for i, k in enumerate(red.names) if k.startswith("st_blocks.1.")]
    assert max(second) < min(first) and max(first) == len(red.names) - 1
    assert red.n_buckets == 2 and red.bucket_bounds[0][1] == red.offsets[first[0]] == red.bucket_bounds[1][0]
    assert not (set(red.names) & dead)
    used = torch.zeros_like(red.flat, dtype=torch.bool)
    for o, sz in zip(red.offsets, red.sizes):
        assert o % 64 == 0                                # slots are 256-byte aligned; the padding between them is zero
        used[o:o + sz] = True
    assert float(red.flat[used].min()) == 2.0 and float(red.flat[~used].abs().max() if (~used).any() else 0.0) == 0.0
    assert all(p.grad.data_ptr() == red.flat.data_ptr() + 4 * o
                                                for p, o in zip(red.live, red.offsets))
    from stgcn_b200.layers import _grad_like
    p0 = red.live[0]
    p0.grad = None
    v = _grad_like(p0, True)                         # the buffer the backward kernels would write
    assert v.data_ptr() == red.flat.data_ptr() + 4 * red.offsets[0] and v.shape == p0.shape


def test_shard_batch():
    from stgcn_b200.dist import shard_batch
    import pytest
    assert [shard_batch(8, r, 4) for r in range(4)] == [slice(0, 2), slice(2, 4), slice(4, 6), slice(6, 8)]
    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)
