"""2-GPU NCCL test of the data-parallel step on the PRODUCT layers (skipped with fewer than 2 GPUs): every rank runs its
batch shard through the captured step (stgcn_b200.graph.GraphedStep) with the flat-buffer reducer -- gradients written by
the backward kernels straight into the flat buffer, ncclAvg all-reduce of it behind every replay -- and the averaged
gradients must equal the full-batch gradients (MSE is a mean over equal shards, LayerNorm is per sample: SURVEY.md §8e)."""
import os
import sys
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, init_file, out_file, precision):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import stgcn_b200
    from conftest import load_gso, rel_l2
    from stgcn_b200.dist import FlatGradAllReducer, shard_batch
    from stgcn_b200.graph import GraphedStep
    from stgcn_b200.synthetic import build_model
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world, device_id=dev)
    stgcn_b200.set_precision(precision)
    gso = load_gso("pemsd7m", "cheb")
    n = gso.shape[0]
    blocks = [[1], [64, 16, 64], [64, 16, 64], [128, 128], [1]]
    B = 16
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(B, 1, 12, n, generator=gen)
    y = torch.randn(B, n, generator=gen)
    model = build_model(gso, "cheb_graph_conv", 3, blocks, dev, seed=7)        # same seed -> replicated parameters
    model.train()
    sl = shard_batch(B, rank, world)
    Bs = sl.stop - sl.start
    red = FlatGradAllReducer(model)
    step = GraphedStep(model, (Bs, 1, 12, n), (Bs, n), device=dev, warmup=2, reducer=red)
    assert red.n_buckets == 2
    for _ in range(2):                       # replays are repeatable
        step(x[sl].to(dev), y[sl].to(dev))
    torch.cuda.synchronize()
    got = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    assert all(p.grad.data_ptr() == red.flat.data_ptr() + 4 * off for p, off in zip(red.live, red.offsets))
    step.close()
    red.unbind()
    if rank == 0:
        model.zero_grad(set_to_none=True)
        torch.nn.functional.mse_loss(model(x.to(dev)).view(B, -1).float(), y.to(dev)).backward()
        torch.cuda.synchronize()
        worst = max(rel_l2(got[k].cpu(), p.grad.cpu()) for k, p in model.named_parameters() if p.grad is not None)
        torch.save({"worst": worst, "n_live": len(got)}, out_file)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 5e-2)])
def test_nccl_averaged_shard_grads_equal_full_batch(precision, tol, cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as td:
        init_file, out_file = os.path.join(td, "init"), os.path.join(td, "out.pt")
        mp.spawn(_worker, args=(2, init_file, out_file, precision), nprocs=2, join=True)
        res = torch.load(out_file)
    assert res["n_live"] == 28
    assert res["worst"] < tol, res
