"""tf32x3 parity mode (csrc/umma_x3.cuh): the fp32 chain with every GEMM on tcgen05 tensor cores (kind::tf32, 3xTF32
operand splitting).  tests/test_gpu_parity.py runs the whole parity suite in this mode; here: proof that the tensor-core
kernel is what actually runs, and layer-level checks at the real graph sizes (multi-tile, ragged tiles, K split over
CTAs, 256-wide tiles) against an fp64 oracle with the north-star tolerance."""
import pytest
import torch

from conftest import load_gso, rel_l2
from oracle import stgcn_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3
TIGHT = 2e-5          # 22 operand bits + fp32 accumulation: outputs agree with fp64 to ~1e-6


@pytest.fixture(autouse=True)
def _x3_mode():
    import stgcn_b200
    stgcn_b200.set_precision("tf32x3")
    yield
    stgcn_b200.set_precision("fp32")


def _d(t):
    return t.detach().double()


def test_x3_kernels_are_what_runs(cuda_device):
    """Per-kernel event profile of one temporal conv + one graph conv layer, forward and backward: every GEMM launch is
    umma_x3_kernel, none is a CUDA-core GEMM."""
    from stgcn_b200 import layers, _lib as L
    dev = cuda_device
    gen = torch.Generator().manual_seed(0)
    tc = layers.TemporalConvLayer(3, 64, 64, 50, "glu").to(dev)
    gc = layers.GraphConvLayer("cheb_graph_conv", 64, 16, 3, O.synthetic_gso(50, seed=1).to(dev), True).to(dev)
    x = torch.randn(2, 64, 8, 50, generator=gen).to(dev).requires_grad_(True)
    L.profile_begin()
    y = gc(tc(x))
    y.sum().backward()
    prof = L.profile_end()
    names = " ".join(prof)
    assert "umma_x3_kernel<TAP>" in names and "umma_x3_kernel<GSO>" in names and "umma_x3_kernel<WGRAD>" in names
    for k in prof:
        assert "tapgemm_kernel" not in k and "gso_kernel" not in k and "wgrad_kernel" not in k and "wgrad_skinny" not in k, k


@pytest.mark.parametrize("c_in,c_out,kt,T,N,B,act", [(64, 64, 3, 8, 228, 6, "glu"), (16, 64, 3, 10, 228, 5, "glu"),
                                                      (64, 128, 4, 4, 325, 7, "glu"), (128, 64, 3, 7, 207, 3, "gtu"),
                                                      (64, 256, 2, 5, 130, 4, "relu"), (8, 24, 3, 6, 41, 3, "silu")])
def test_x3_temporal_conv_full_size(c_in, c_out, kt, T, N, B, act, cuda_device):
    from stgcn_b200 import layers
    dev = cuda_device
    gen = torch.Generator().manual_seed(c_in + c_out + kt + N)
    p = {}
    O._tconv_params(p, "t.", kt, c_in, c_out, act, gen)
    layer = layers.TemporalConvLayer(kt, c_in, c_out, N, act).to(dev)
    layer.load_state_dict({k[2:]: v for k, v in p.items()}, strict=True)
    x = torch.randn(B, c_in, T, N, generator=gen)
    xg = x.to(dev).requires_grad_(True)
    y = layer(xg)
    pr = {k: v.double().requires_grad_(True) for k, v in p.items()}
    xr = x.double().requires_grad_(True)
    yr = O.temporal_gated_conv(xr, pr, "t.", kt, c_out, act)
    assert y.dtype == torch.float32 and tuple(y.shape) == tuple(yr.shape)
    assert rel_l2(y.cpu(), yr) < TIGHT
    dy = torch.randn(yr.shape, generator=gen)
    y.backward(dy.to(dev))
    yr.backward(dy.double())
    assert rel_l2(xg.grad.cpu(), xr.grad) < 1e-4
    named = dict(layer.named_parameters())
    for k, v in pr.items():
        if v.grad is not None:
            assert rel_l2(named[k[2:]].grad.cpu(), v.grad) < 1e-4, k


@pytest.mark.parametrize("kind,ks", [("cheb_graph_conv", 3), ("cheb_graph_conv", 5), ("graph_conv", 3)])
@pytest.mark.parametrize("c_in,c,N,B,T", [(64, 16, 228, 3, 10), (64, 16, 325, 2, 6), (16, 16, 207, 3, 5),
                                          (64, 64, 300, 1, 3), (12, 8, 41, 2, 7)])
def test_x3_graph_conv_full_size(kind, ks, c_in, c, N, B, T, cuda_device):
    """Non-symmetric operator (catches a missing transpose in the adjoint), N not a multiple of the 128-row tile or of
    the 32-element K chunk, 16- and 64-channel planes."""
    from stgcn_b200 import layers
    dev = cuda_device
    gen = torch.Generator().manual_seed(ks * 131 + c_in + N)
    torch.manual_seed(ks * 131 + c_in + N)
    a = torch.randn(N, N, generator=gen)
    gso = (a / torch.linalg.matrix_norm(a, ord=2)).float()
    layer = layers.GraphConvLayer(kind, c_in, c, ks, gso.to(dev), True).to(dev)
    p = {"g." + k: v.detach().cpu().clone() for k, v in layer.state_dict().items()}
    x = torch.randn(B, c_in, T, N, generator=gen)
    xg = x.to(dev).requires_grad_(True)
    y = layer(xg, _relu=1)
    pr = {k: v.double().requires_grad_(True) for k, v in p.items()}
    xr = x.double().requires_grad_(True)
    yr = torch.relu(O.graph_conv_layer(xr, pr, "g.", gso.double(), c, kind))
    assert rel_l2(y.cpu(), yr) < TIGHT
    dy = torch.randn(yr.shape, generator=gen)
    y.backward(dy.to(dev))
    yr.backward(dy.double())
    assert rel_l2(xg.grad.cpu(), xr.grad) < TOL
    named = dict(layer.named_parameters())
    for k, v in pr.items():
        if v.grad is not None:
            assert rel_l2(named[k[2:]].grad.cpu(), v.grad) < TOL, k


def test_x3_weight_gradient_split_over_ctas(cuda_device):
    """Enough rows that the weight-gradient contraction is split over many CTAs (partials + reduction): B*T*N = 233 472
    rows for the 64 -> 128-wide conv and its 193-row (3 taps x 64 channels + bias) gradient."""
    from stgcn_b200 import layers
    dev = cuda_device
    gen = torch.Generator().manual_seed(3)
    B, T, N, c = 128, 8, 228, 64
    p = {}
    O._tconv_params(p, "t.", 3, c, c, "glu", gen)
    layer = layers.TemporalConvLayer(3, c, c, N, "glu").to(dev)
    layer.load_state_dict({k[2:]: v for k, v in p.items()}, strict=True)
    x = torch.randn(B, c, T, N, generator=gen)
    dy = torch.randn(B, c, T - 2, N, generator=gen)
    xg = x.to(dev).requires_grad_(True)
    layer(xg).backward(dy.to(dev))
    pr = {k: v.double().requires_grad_(True) for k, v in p.items()}
    xr = x.double().requires_grad_(True)
    O.temporal_gated_conv(xr, pr, "t.", 3, c, "glu").backward(dy.double())
    named = dict(layer.named_parameters())
    assert rel_l2(named["causal_conv.weight"].grad.cpu(), pr["t.causal_conv.weight"].grad) < 1e-4
    assert rel_l2(named["causal_conv.bias"].grad.cpu(), pr["t.causal_conv.bias"].grad) < 1e-4
    assert rel_l2(xg.grad.cpu(), xr.grad) < 1e-4


def test_x3_full_model_b256_additivity(cuda_device):
    """BASELINE's B = 256 in the tensor-core parity mode: shard gradients add up to the batch gradient (size-independent
    property; the fp64 oracle at this size would take minutes)."""
    from types import SimpleNamespace
    from stgcn_b200 import models
    dev = cuda_device
    gso = load_gso("pemsd7m", "cheb")
    n = gso.shape[0]
    blocks = [[1], [64, 16, 64], [64, 16, 64], [128, 128], [1]]
    args = SimpleNamespace(Kt=3, Ks=3, act_func="glu", graph_conv_type="cheb_graph_conv", gso=gso.to(dev),
                           enable_bias=True, droprate=0.0, n_his=12)
    model = models.STGCNChebGraphConv(args, blocks, n).to(dev)
    model.load_state_dict(O.init_params(blocks=blocks, kt=3, ks=3, n_his=12, n_vertex=n, seed=4), strict=True)
    B = 256
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(B, 1, 12, n, generator=gen).to(dev)
    y = torch.randn(B, n, generator=gen).to(dev)

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
