"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the reference-generated
golden vectors.  Tolerance: north_star's 1e-3 relative (per-tensor rel-L2); exact for index work.

Every test runs in BOTH parity modes: ``fp32`` (CUDA-core kernels) and ``tf32x3`` (the same fp32 chain with every GEMM
on tcgen05 tensor cores, 3xTF32 operand splitting -- csrc/umma_x3.cuh).  Same tolerances for both."""
import math
from types import SimpleNamespace

import pytest
import torch

from conftest import GoldenCase, golden_case_names, load_gso, rel_l2
from oracle import stgcn_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-3          # the parity gate of BASELINE.json's north_star
TIGHT = 2e-4        # what both parity modes are expected to achieve on outputs


@pytest.fixture(autouse=True, params=["fp32", "tf32x3"])
def parity_mode(request):
    import stgcn_b200
    stgcn_b200.set_precision(request.param)
    yield request.param
    stgcn_b200.set_precision("fp32")


def _model_from_cfg(cfg, gso, dev, droprate=0.0):
    from stgcn_b200 import models
    args = SimpleNamespace(Kt=cfg["Kt"], Ks=cfg["Ks"], act_func=cfg["act"], graph_conv_type=cfg["kind"],
                           gso=gso.to(dev), enable_bias=cfg["bias"], droprate=droprate, n_his=cfg["n_his"])
    cls = models.STGCNChebGraphConv if cfg["kind"] == "cheb_graph_conv" else models.STGCNGraphConv
    return cls(args, cfg["blocks"], cfg["n"]).to(dev)


@pytest.mark.parametrize("name", golden_case_names())
def test_model_matches_reference_golden(name, cuda_device):
    g = GoldenCase(name)
    dev = cuda_device
    model = _model_from_cfg(g.cfg, g.gso, dev)
    model.load_state_dict(g.params, strict=True)
    model.train()
    x = g.x.to(dev).requires_grad_(True)
    out = model(x)
    assert tuple(out.shape) == tuple(g.out.shape)
    assert rel_l2(out.cpu(), g.out) < TIGHT
    b0 = model.st_blocks[0](g.x.to(dev))
    assert tuple(b0.shape) == tuple(g.block0_out.shape)
    assert rel_l2(b0.cpu(), g.block0_out) < TIGHT
    B = x.shape[0]
    loss = torch.nn.functional.mse_loss(out.reshape(B, -1), g.y.to(dev))
    assert abs(loss.item() - g.loss) < TOL * max(1.0, abs(g.loss))
    loss.backward()
    assert rel_l2(x.grad.cpu(), g.dx) < TOL
    named = dict(model.named_parameters())
    for k, gref in g.grads.items():
        assert named[k].grad is not None, k
        assert rel_l2(named[k].grad.cpu(), gref) < TOL, k
    for k, p in named.items():          # dead align convs get no gradient, as in the reference
        if k not in g.grads:
            assert p.grad is None, k


def _rand_tconv(c_in, c_out, kt, act, gen):
    p = {}
    O._tconv_params(p, "t.", kt, c_in, c_out, act, gen)
    return p


@pytest.mark.parametrize("act", ["glu", "gtu", "relu", "silu"])
@pytest.mark.parametrize("c_in,c_out", [(1, 64), (64, 64), (16, 64), (12, 8), (5, 7), (3, 3)])
@pytest.mark.parametrize("kt", [2, 3])
def test_temporal_conv_layer(act, c_in, c_out, kt, cuda_device):
    from stgcn_b200 import layers
    dev = cuda_device
    gen = torch.Generator().manual_seed(c_in * 131 + c_out * 7 + kt)
    B, T, N = 3, 9, 29
    p = _rand_tconv(c_in, c_out, kt, act, gen)
    layer = layers.TemporalConvLayer(kt, c_in, c_out, N, act).to(dev)
    layer.load_state_dict({k[2:]: v for k, v in p.items()}, strict=True)
    x = torch.randn(B, c_in, T, N, generator=gen)
    xg = x.to(dev).requires_grad_(True)
    y = layer(xg)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    xr = x.clone().requires_grad_(True)
    yr = O.temporal_gated_conv(xr, pr, "t.", kt, c_out, act)
    assert tuple(y.shape) == tuple(yr.shape)
    assert rel_l2(y.cpu(), yr) < TIGHT
    dy = torch.randn(yr.shape, generator=gen)
    y.backward(dy.to(dev))
    yr.backward(dy)
    assert rel_l2(xg.grad.cpu(), xr.grad) < TOL
    named = dict(layer.named_parameters())
    for k, v in pr.items():
        if v.grad is None:
            assert named[k[2:]].grad is None, k
        else:
            assert rel_l2(named[k[2:]].grad.cpu(), v.grad) < TOL, k


@pytest.mark.parametrize("kind,ks", [("cheb_graph_conv", 1), ("cheb_graph_conv", 2), ("cheb_graph_conv", 3),
                                      ("cheb_graph_conv", 5), ("graph_conv", 3)])
@pytest.mark.parametrize("c_in,c_out", [(64, 16), (8, 8), (4, 8), (6, 5)])
@pytest.mark.parametrize("bias", [True, False])
def test_graph_conv_layer(kind, ks, c_in, c_out, bias, cuda_device):
    from stgcn_b200 import layers
    dev = cuda_device
    gen = torch.Generator().manual_seed(ks * 17 + c_in * 3 + c_out)
    B, T, N = 2, 5, 41
    a = torch.randn(N, N, generator=gen)
    gso = (a / torch.linalg.matrix_norm(a, ord=2)).float()          # deliberately non-symmetric
    layer = layers.GraphConvLayer(kind, c_in, c_out, ks, gso.to(dev), bias).to(dev)
    p = {"g." + k: v.detach().cpu().clone() for k, v in layer.state_dict().items()}
    x = torch.randn(B, c_in, T, N, generator=gen)
    xg = x.to(dev).requires_grad_(True)
    y = layer(xg)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    xr = x.clone().requires_grad_(True)
    yr = O.graph_conv_layer(xr, pr, "g.", gso, c_out, kind)
    assert tuple(y.shape) == tuple(yr.shape)
    assert rel_l2(y.cpu(), yr) < TIGHT
    dy = torch.randn(yr.shape, generator=gen)
    y.backward(dy.to(dev))
    yr.backward(dy)
    assert rel_l2(xg.grad.cpu(), xr.grad) < TOL
    named = dict(layer.named_parameters())
    for k, v in pr.items():
        if v.grad is None:
            assert named[k[2:]].grad is None, k
        else:
            assert rel_l2(named[k[2:]].grad.cpu(), v.grad) < TOL, k


@pytest.mark.parametrize("ks", [1, 3])
def test_bare_cheb_and_gcn_conv(ks, cuda_device):
    from stgcn_b200 import layers
    dev = cuda_device
    gen = torch.Generator().manual_seed(ks)
    B, T, N, Cc = 2, 4, 23, 16
    gso = O.synthetic_gso(N, seed=ks)
    x = torch.randn(B, Cc, T, N, generator=gen)
    cheb = layers.ChebGraphConv(Cc, Cc, ks, gso.to(dev), True).to(dev)
    y = cheb(x.to(dev))
    yr = O.cheb_graph_conv(x, gso, cheb.weight.detach().cpu(), cheb.bias.detach().cpu())
    assert tuple(y.shape) == (B, T, N, Cc) and rel_l2(y.cpu(), yr) < TIGHT
    gcn = layers.GraphConv(Cc, Cc, gso.to(dev), False).to(dev)
    y = gcn(x.to(dev))
    yr = O.first_order_graph_conv(x, gso, gcn.weight.detach().cpu(), None)
    assert tuple(y.shape) == (B, T, N, Cc) and rel_l2(y.cpu(), yr) < TIGHT


def test_align_and_causal_conv(cuda_device):
    from stgcn_b200 import layers
    dev = cuda_device
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 12, 6, 19, generator=gen)
    al = layers.Align(12, 5).to(dev)
    y = al(x.to(dev))
    yr = torch.nn.functional.conv2d(x, al.align_conv.weight.detach().cpu(), al.align_conv.bias.detach().cpu())
    assert rel_l2(y.cpu(), yr) < TIGHT
    # growing / equal: pure index work, bit exact
    xi = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).reshape(2, 3, 4, 5).to(dev)
    up = layers.Align(3, 7).to(dev)(xi)
    assert up.shape == (2, 7, 4, 5) and torch.equal(up[:, :3], xi) and torch.count_nonzero(up[:, 3:]) == 0
    assert layers.Align(3, 3).to(dev)(xi) is xi
    cc = layers.CausalConv2d(12, 9, (3, 1)).to(dev)
    y = cc(x.to(dev))
    yr = torch.nn.functional.conv2d(x, cc.weight.detach().cpu(), cc.bias.detach().cpu())
    assert tuple(y.shape) == tuple(yr.shape) and rel_l2(y.cpu(), yr) < TIGHT


def test_index_work_bit_exact(cuda_device):
    """With zero conv weights GLU returns exactly 0.5 * zero-padded x[:, :, Kt-1:]: pins the slice/pad/permute
    index arithmetic bit-exactly."""
    from stgcn_b200 import layers
    dev = cuda_device
    layer = layers.TemporalConvLayer(3, 3, 6, 11, "glu").to(dev)
    with torch.no_grad():
        layer.causal_conv.weight.zero_()
        layer.causal_conv.bias.zero_()
    x = torch.arange(2 * 3 * 7 * 11, dtype=torch.float32).reshape(2, 3, 7, 11)
    y = layer(x.to(dev)).cpu()
    assert y.shape == (2, 6, 5, 11)
    assert torch.equal(y[:, :3], 0.5 * x[:, :, 2:]) and torch.count_nonzero(y[:, 3:]) == 0
    # graph conv with identity operator, Ks=1, W=I, b=0 -> relu-less output is exactly 2x
    gl = layers.GraphConvLayer("cheb_graph_conv", 4, 4, 1, torch.eye(11).to(dev), True).to(dev)
    with torch.no_grad():
        gl.cheb_graph_conv.weight.copy_(torch.eye(4).reshape(1, 4, 4))
        gl.cheb_graph_conv.bias.zero_()
    xi = torch.arange(2 * 4 * 3 * 11, dtype=torch.float32).reshape(2, 4, 3, 11)
    assert torch.equal(gl(xi.to(dev)).cpu(), 2 * xi)


def test_layer_norm_and_dropout_op(cuda_device):
    from stgcn_b200 import layers
    dev = cuda_device
    gen = torch.Generator().manual_seed(11)
    B, T, N, Cc = 3, 4, 37, 24
    x = torch.randn(B, T, N, Cc, generator=gen) * 3 + 1
    w = torch.randn(N, Cc, generator=gen)
    b = torch.randn(N, Cc, generator=gen)
    xg = x.to(dev).requires_grad_(True)
    wg, bg = w.to(dev).requires_grad_(True), b.to(dev).requires_grad_(True)
    y = layers._LnormFn.apply(xg, (B, T, N, Cc, False, 0.0, 1e-12), wg, bg)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (N, Cc), wr, br, 1e-12)
    assert rel_l2(y.cpu(), yr) < TIGHT
    dy = torch.randn(yr.shape, generator=gen)
    y.backward(dy.to(dev)); yr.backward(dy)
    assert rel_l2(xg.grad.cpu(), xr.grad) < TOL
    assert rel_l2(wg.grad.cpu(), wr.grad) < TOL and rel_l2(bg.grad.cpu(), br.grad) < TOL
    # dropout: kept entries are the eval output / (1-p); the backward uses the same mask
    p = 0.5
    xg2 = x.to(dev).requires_grad_(True)
    torch.manual_seed(1234)
    yd = layers._LnormFn.apply(xg2, (B, T, N, Cc, True, p, 1e-12), wg.detach(), bg.detach())
    kept = yd != 0
    frac = kept.float().mean().item()
    assert abs(frac - (1 - p)) < 0.02
    assert torch.allclose(yd[kept], (y.detach() / (1 - p))[kept], rtol=1e-5, atol=1e-6)
    yd.backward(torch.ones_like(yd))
    # input gradient of a fully dropped group element pattern: compare against autograd through a masked LN
    mask = kept.float() / (1 - p)
    xr2 = x.clone().requires_grad_(True)
    (torch.nn.functional.layer_norm(xr2, (N, Cc), w, b, 1e-12) * mask.cpu()).sum().backward()
    assert rel_l2(xg2.grad.cpu(), xr2.grad) < TOL


@pytest.mark.parametrize("dataset,kind,B", [("pemsd7m", "cheb_graph_conv", 32), ("metrla", "graph_conv", 16),
                                            ("pemsbay", "cheb_graph_conv", 8)])
def test_full_size_model_vs_oracle(dataset, kind, B, cuda_device):
    """BASELINE.json configs at their real graph sizes (small batch so the CPU oracle finishes in seconds)."""
    dev = cuda_device
    gso = load_gso(dataset, "cheb" if kind == "cheb_graph_conv" else "gcn")
    n = gso.shape[0]
    blocks = [[1], [64, 16, 64], [64, 16, 64], [128, 128], [1]]
    cfg = dict(Kt=3, Ks=3, act="glu", kind=kind, bias=True, n_his=12, blocks=blocks, n=n)
    params = O.init_params(blocks=blocks, kt=3, ks=3, n_his=12, n_vertex=n, kind=kind, seed=2)
    model = _model_from_cfg(cfg, gso, dev)
    model.load_state_dict(params, strict=True)
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(B, 1, 12, n, generator=gen)
    y = torch.randn(B, n, generator=gen)
    out = model(x.to(dev))
    loss = torch.nn.functional.mse_loss(out.view(B, -1), y.to(dev))
    loss.backward()
    # truth: fp64 oracle; floor: the fp32 oracle's own distance to it (SURVEY.md §7 hard part 1)
    p64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
    l64 = O.mse_step(x.double(), y.double(), p64, gso.double(), blocks=blocks, kt=3, n_his=12, kind=kind)
    l64.backward()
    p32 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    l32 = O.mse_step(x, y, p32, gso, blocks=blocks, kt=3, n_his=12, kind=kind)
    l32.backward()
    assert abs(loss.item() - l64.item()) < TOL * abs(l64.item())
    named = dict(model.named_parameters())
    for k, v in p64.items():
        if v.grad is None:
            continue
        ours = rel_l2(named[k].grad.cpu(), v.grad)
        floor = rel_l2(p32[k].grad, v.grad)
        assert ours < max(TOL, 3 * floor), (k, ours, floor)


def test_batch_independence_and_gradient_additivity(cuda_device):
    """Size-independent properties at BASELINE's B=256: a sample's output does not depend on its batch mates
    (bit exact: same arithmetic order per row), and the batch gradient is the sum of shard gradients."""
    dev = cuda_device
    gso = load_gso("pemsd7m", "cheb")
    n = gso.shape[0]
    blocks = [[1], [64, 16, 64], [64, 16, 64], [128, 128], [1]]
    cfg = dict(Kt=3, Ks=3, act="glu", kind="cheb_graph_conv", bias=True, n_his=12, blocks=blocks, n=n)
    model = _model_from_cfg(cfg, gso, dev)
    model.load_state_dict(O.init_params(blocks=blocks, kt=3, ks=3, n_his=12, n_vertex=n, seed=4), strict=True)
    B = 256
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(B, 1, 12, n, generator=gen).to(dev)
    y = torch.randn(B, n, generator=gen).to(dev)
    with torch.no_grad():
        full = model(x)
        part = model(x[100:103])
    assert torch.equal(full[100:103], part)

    def grads(xs, ys, scale):
        model.zero_grad(set_to_none=True)
        out = model(xs).view(xs.shape[0], -1)
        (torch.nn.functional.mse_loss(out, ys, reduction="sum") * scale).backward()
        return {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}

    s = 1.0 / (B * n)
    g_all = grads(x, y, s)
    g_a = grads(x[:128], y[:128], s)
    g_b = grads(x[128:], y[128:], s)
    for k in g_all:
        assert rel_l2(g_a[k] + g_b[k], g_all[k]) < 1e-4, k


def test_dropout_training_mode_block(cuda_device):
    from stgcn_b200 import layers
    dev = cuda_device
    n = 31
    gso = O.synthetic_gso(n, seed=9).to(dev)
    blk = layers.STConvBlock(3, 3, n, 1, [16, 8, 16], "glu", "cheb_graph_conv", gso, True, 0.5).to(dev)
    x = torch.randn(8, 1, 12, n, device=dev)
    blk.eval()
    y_eval = blk(x)
    blk.train()
    torch.manual_seed(7)
    xt = x.clone().requires_grad_(True)
    y_tr = blk(xt)
    kept = y_tr != 0
    assert abs(kept.float().mean().item() - 0.5) < 0.02
    assert torch.allclose(y_tr[kept], 2 * y_eval[kept], rtol=1e-5, atol=1e-6)
    torch.manual_seed(7)
    y_tr2 = blk(x)          # different call -> different mask (counter-based seeds)
    assert not torch.equal(y_tr2 != 0, kept)
    y_tr.sum().backward()
    assert torch.isfinite(xt.grad).all()


def test_error_behaviour(cuda_device):
    from stgcn_b200 import layers, StgcnError
    dev = cuda_device
    x = torch.randn(2, 4, 6, 9, device=dev)
    with pytest.raises(NotImplementedError):
        layers.TemporalConvLayer(3, 4, 4, 9, "tanh").to(dev)(x)
    with pytest.raises(ValueError):
        layers.GraphConvLayer("cheb_graph_conv", 4, 4, 0, torch.eye(9), True).to(dev)(x)
    with pytest.raises(RuntimeError):
        layers.TemporalConvLayer(3, 4, 4, 9, "glu").to(dev)(x.cpu())
    with pytest.raises(StgcnError):      # time axis shorter than the kernel, like conv2d's own error
        layers.TemporalConvLayer(7, 4, 4, 9, "glu").to(dev)(x)


def test_mse_helper(cuda_device):
    import ctypes as C
    from stgcn_b200 import _lib as L
    dev = cuda_device
    pred = torch.randn(1000, device=dev)
    tgt = torch.randn(1000, device=dev)
    loss = torch.zeros(1, device=dev)
    dpred = torch.empty_like(pred)
    L.check(L.lib().stgcn_mse_fwd_bwd(pred.data_ptr(), tgt.data_ptr(), 1000, 1.0, loss.data_ptr(), dpred.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream))
    ref = torch.nn.functional.mse_loss(pred, tgt)
    assert abs(loss.item() - ref.item()) < 1e-5
    assert torch.allclose(dpred, 2 * (pred - tgt) / 1000, rtol=1e-5, atol=1e-7)
