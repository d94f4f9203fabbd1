"""bf16 (throughput) mode: activations stored in bf16, fp32 accumulation.  This mode is NOT the 1e-3 parity
gate (that is the fp32 mode, tests/test_gpu_parity.py); its documented tolerance follows SURVEY.md §7 hard
part 1 (bf16 operand rounding: ~4e-3 on outputs, ~1e-2..6e-2 on gradients vs an fp64 truth)."""
from types import SimpleNamespace

import pytest
import torch

from conftest import GoldenCase, golden_case_names, load_gso, rel_l2
from oracle import stgcn_oracle as O

pytestmark = pytest.mark.gpu

OUT_TOL = 3e-2
GRAD_TOL = 1e-1


@pytest.fixture(autouse=True)
def _bf16_mode():
    import stgcn_b200
    stgcn_b200.set_precision("bf16")
    yield
    stgcn_b200.set_precision("fp32")


def _model(cfg, gso, dev):
    from stgcn_b200 import models
    args = SimpleNamespace(Kt=cfg["Kt"], Ks=cfg["Ks"], act_func=cfg["act"], graph_conv_type=cfg["kind"],
                           gso=gso.to(dev), enable_bias=cfg["bias"], droprate=0.0, n_his=cfg["n_his"])
    cls = models.STGCNChebGraphConv if cfg["kind"] == "cheb_graph_conv" else models.STGCNGraphConv
    return cls(args, cfg["blocks"], cfg["n"]).to(dev)


def _emulated_errors(g, draws=6):
    """Per-tensor error of the bf16-storage error model (tests/bf16_emulation.py): max over a few noise draws (the
    input is perturbed by 1% so the rounding pattern is re-drawn; the fp32 oracle on the same input is the truth).
    The distribution is heavy-tailed for tiny networks -- a single ReLU/GTU mask flip moves a small-batch gradient by
    tens of percent -- so one draw is not a bound."""
    import bf16_emulation as E
    worst = {}
    for trial in range(draws):
        gen = torch.Generator().manual_seed(trial)
        x0 = g.x if trial == 0 else g.x * (1 + 1e-2 * torch.randn(g.x.shape, generator=gen))
        res = {}
        for mode in ("ref", "emu"):
            params = {k: v.clone().requires_grad_(True) for k, v in g.params.items()}
            x = x0.clone().requires_grad_(True)
            fn = O.stgcn_forward if mode == "ref" else E.forward
            out = fn(x, params, g.gso, **g.model_cfg())
            torch.nn.functional.mse_loss(out.reshape(x.shape[0], -1), g.y).backward()
            res[mode] = {"out": out.detach(), "dx": x.grad, **{"g:" + k: params[k].grad for k in g.grads}}
        for k in res["ref"]:
            worst[k] = max(worst.get(k, 0.0), rel_l2(res["emu"][k], res["ref"][k]))
    return worst


@pytest.mark.parametrize("name", golden_case_names())
def test_bf16_model_close_to_reference_golden(name, cuda_device):
    """bf16 mode vs the reference's golden vectors.  Bound: outputs 3e-2; every gradient tensor within
    max(GRAD_TOL, 2x the bf16-storage error model's worst draw) -- i.e. no worse than what storing activations in bf16 costs."""
    g = GoldenCase(name)
    dev = cuda_device
    model = _model(g.cfg, g.gso, dev)
    model.load_state_dict(g.params, strict=True)
    model.train()
    x = g.x.to(dev).requires_grad_(True)
    out = model(x).float()
    assert tuple(out.shape) == tuple(g.out.shape)
    B = x.shape[0]
    loss = torch.nn.functional.mse_loss(out.reshape(B, -1), g.y.to(dev))
    loss.backward()
    named = dict(model.named_parameters())
    errs = {"out": rel_l2(out.cpu(), g.out), "dx": rel_l2(x.grad.cpu(), g.dx)}
    for k, gref in g.grads.items():
        assert named[k].grad is not None, k
        errs["g:" + k] = rel_l2(named[k].grad.cpu(), gref)
    model_err = _emulated_errors(g)
    worst_model = max(v for k, v in model_err.items() if k != "out")
    bad = {}
    for k, v in errs.items():
        bound = OUT_TOL if k == "out" else max(GRAD_TOL, 2.0 * model_err[k], 1.0 * worst_model)
        if v > bound:
            bad[k] = (round(v, 4), round(bound, 4))
    assert not bad, bad
    assert abs(loss.item() - g.loss) < OUT_TOL * max(1.0, abs(g.loss))


@pytest.mark.parametrize("dataset,kind,B", [("pemsd7m", "cheb_graph_conv", 16), ("metrla", "graph_conv", 8),
                                            ("pemsbay", "cheb_graph_conv", 8)])
def test_bf16_full_size_model(dataset, kind, B, cuda_device):
    dev = cuda_device
    gso = load_gso(dataset, "cheb" if kind == "cheb_graph_conv" else "gcn")
    n = gso.shape[0]
    blocks = [[1], [64, 16, 64], [64, 16, 64], [128, 128], [1]]
    cfg = dict(Kt=3, Ks=3, act="glu", kind=kind, bias=True, n_his=12, blocks=blocks, n=n)
    params = O.init_params(blocks=blocks, kt=3, ks=3, n_his=12, n_vertex=n, kind=kind, seed=2)
    model = _model(cfg, gso, dev)
    model.load_state_dict(params, strict=True)
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(B, 1, 12, n, generator=gen)
    y = torch.randn(B, n, generator=gen)
    out = model(x.to(dev))
    assert out.dtype == torch.float32           # the model output (loss input) stays fp32
    loss = torch.nn.functional.mse_loss(out.view(B, -1), y.to(dev))
    loss.backward()
    p64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
    l64 = O.mse_step(x.double(), y.double(), p64, gso.double(), blocks=blocks, kt=3, n_his=12, kind=kind)
    l64.backward()
    assert abs(loss.item() - l64.item()) < OUT_TOL * abs(l64.item())
    named = dict(model.named_parameters())
    errs = {k: rel_l2(named[k].grad.cpu(), v.grad) for k, v in p64.items() if v.grad is not None}
    assert max(errs.values()) < 1.5 * GRAD_TOL, max(errs.items(), key=lambda kv: kv[1])   # small B: bias sums are noisy
    assert sorted(errs.values())[len(errs) // 2] < 8e-2


@pytest.mark.parametrize("c_in,c_out,kt,T", [(64, 64, 3, 8), (16, 64, 3, 10), (64, 128, 4, 4), (32, 32, 2, 6),
                                              (128, 64, 3, 7)])
@pytest.mark.parametrize("act", ["glu", "relu"])
def test_bf16_tcgen05_temporal_conv(c_in, c_out, kt, T, act, cuda_device):
    """Shapes served by the tcgen05 tap-GEMM kernel (csrc/umma_tap.cuh).  Oracle evaluated on the same
    bf16-rounded inputs/weights, so the remaining error is accumulation order + bf16 output rounding."""
    from stgcn_b200 import layers
    dev = cuda_device
    gen = torch.Generator().manual_seed(c_in + c_out + kt)
    B, N = 3, 228
    p = {}
    O._tconv_params(p, "t.", kt, c_in, c_out, act, gen)
    layer = layers.TemporalConvLayer(kt, c_in, c_out, N, act).to(dev)
    layer.load_state_dict({k[2:]: v for k, v in p.items()}, strict=True)
    x = torch.randn(B, c_in, T, N, generator=gen)
    xg = x.to(dev).requires_grad_(True)
    y = layer(xg)
    assert y.dtype == torch.bfloat16
    rb = lambda t: t.bfloat16().float()
    pr = {k: rb(v).requires_grad_(True) if "causal_conv.weight" in k or "align" in k else v.clone().requires_grad_(True)
          for k, v in p.items()}
    xr = rb(x).requires_grad_(True)
    yr = O.temporal_gated_conv(xr, pr, "t.", kt, c_out, act)
    assert tuple(y.shape) == tuple(yr.shape)
    assert rel_l2(y.float().cpu(), yr) < 6e-3
    dy = torch.randn(yr.shape, generator=gen)
    y.backward(dy.to(dev).bfloat16())
    yr.backward(rb(dy))
    # ReLU's mask is recomputed from the bf16-rounded pre-activation: elements within rounding distance of 0 flip
    tol = 2e-2 if act == "glu" else 6e-2
    assert rel_l2(xg.grad.cpu(), xr.grad) < tol
    named = dict(layer.named_parameters())
    for k, v in pr.items():
        if v.grad is not None:
            assert rel_l2(named[k[2:]].grad.cpu(), v.grad) < tol, k


@pytest.mark.parametrize("kind,ks", [("cheb_graph_conv", 2), ("cheb_graph_conv", 3), ("cheb_graph_conv", 5),
                                      ("graph_conv", 3)])
@pytest.mark.parametrize("c_in,N,B,T", [(16, 228, 3, 5), (64, 228, 2, 3), (16, 41, 2, 7), (64, 207, 1, 9),
                                        (16, 130, 4, 3)])
@pytest.mark.parametrize("relu", [0, 1])
def test_bf16_fused_graph_conv_layer(kind, ks, c_in, N, B, T, relu, cuda_device):
    """The fused tcgen05 graph-convolution kernels (csrc/umma_cheb.cuh): Chebyshev recurrence / first-order
    propagation + weight GEMMs + bias + residual (+ ReLU) in one pass, and the adjoint in one pass.  The operator is
    deliberately non-symmetric (catches a missing transpose); ragged last work item (B*T not a multiple of the
    groups per item); N below / above one 128-row tile.  Oracle evaluated on the same bf16-rounded operands."""
    from stgcn_b200 import layers
    dev = cuda_device
    gen = torch.Generator().manual_seed(ks * 131 + c_in + N)
    torch.manual_seed(ks * 131 + c_in + N)          # the layer's parameter init draws from the global generator
    a = torch.randn(N, N, generator=gen)
    gso = (a / torch.linalg.matrix_norm(a, ord=2)).float()
    layer = layers.GraphConvLayer(kind, c_in, 16, ks, gso.to(dev), True).to(dev)
    p = {"g." + k: v.detach().cpu().clone() for k, v in layer.state_dict().items()}
    x = torch.randn(B, c_in, T, N, generator=gen)
    xg = x.to(dev).requires_grad_(True)
    y = layer(xg, _relu=relu)
    assert y.dtype == torch.bfloat16
    rb = lambda t: t.bfloat16().float()
    pr = {k: (rb(v) if v.dim() > 1 else v.clone()).requires_grad_(True) for k, v in p.items()}
    xr = rb(x).requires_grad_(True)
    yr = O.graph_conv_layer(xr, pr, "g.", rb(gso), 16, kind)
    if relu:
        yr = torch.relu(yr)
    assert tuple(y.shape) == tuple(yr.shape)
    assert rel_l2(y.float().cpu(), yr) < 1e-2
    dy = torch.randn(yr.shape, generator=gen)
    y.backward(dy.to(dev).bfloat16())
    yr.backward(rb(dy))
    # ReLU: the mask is recomputed from the bf16-rounded output, so elements within rounding distance of 0 flip
    # (~1% of them); each flip moves a bias-gradient column sum by a whole dy element
    tol = 3e-2 if not relu else 1e-1
    assert rel_l2(xg.grad.cpu(), xr.grad) < tol
    named = dict(layer.named_parameters())
    for k, v in pr.items():
        if v.grad is not None:
            assert rel_l2(named[k[2:]].grad.cpu(), v.grad) < tol, k


@pytest.mark.parametrize("kind,ks,c_in,N,B,T", [("cheb_graph_conv", 3, 64, 325, 2, 6), ("cheb_graph_conv", 3, 16, 325, 3, 4),
                                                  ("graph_conv", 3, 64, 325, 2, 3), ("cheb_graph_conv", 5, 64, 2048, 1, 2),
                                                  ("cheb_graph_conv", 2, 64, 1100, 1, 3)])
@pytest.mark.parametrize("c", [16, 64])
def test_bf16_graph_conv_large_n(kind, ks, c_in, N, B, T, c, cuda_device):
    """Graph sizes the fused TMEM-resident kernel does not take (N = 325: PEMS-BAY, BASELINE configs[3]) and the K-tiled
    node contraction for operators that do not fit shared memory (N = 2048, 64 channels: BASELINE configs[4]).
    Oracle evaluated on the same bf16-rounded operands."""
    from stgcn_b200 import layers
    if c_in < c:
        pytest.skip("the reference never widens in the graph-conv layer")
    dev = cuda_device
    gen = torch.Generator().manual_seed(ks * 7 + c + N)
    torch.manual_seed(ks * 7 + c + N)
    a = torch.randn(N, N, generator=gen)
    gso = (a / torch.linalg.matrix_norm(a, ord=2)).float()
    layer = layers.GraphConvLayer(kind, c_in, c, ks, gso.to(dev), True).to(dev)
    p = {"g." + k: v.detach().cpu().clone() for k, v in layer.state_dict().items()}
    x = torch.randn(B, c_in, T, N, generator=gen)
    xg = x.to(dev).requires_grad_(True)
    y = layer(xg, _relu=0)
    assert y.dtype == torch.bfloat16
    rb = lambda t: t.bfloat16().float()
    pr = {k: (rb(v) if v.dim() > 1 else v.clone()).requires_grad_(True) for k, v in p.items()}
    xr = rb(x).requires_grad_(True)
    yr = O.graph_conv_layer(xr, pr, "g.", rb(gso), c, kind)
    assert rel_l2(y.float().cpu(), yr) < 1.5e-2
    dy = torch.randn(yr.shape, generator=gen)
    y.backward(dy.to(dev).bfloat16())
    yr.backward(rb(dy))
    assert rel_l2(xg.grad.cpu(), xr.grad) < 4e-2
    named = dict(layer.named_parameters())
    for k, v in pr.items():
        if v.grad is not None:
            assert rel_l2(named[k[2:]].grad.cpu(), v.grad) < 4e-2, k


@pytest.mark.parametrize("dataset,kind,B", [("pemsd7m", "cheb_graph_conv", 8), ("metrla", "graph_conv", 4)])
def test_bf16_graphed_step_matches_eager(dataset, kind, B, cuda_device):
    """The whole step (forward + MSE + backward) is captured in ONE CUDA graph although the block-level calls fork helper
    streams for parameter-only work and for the weight-gradient kernels (csrc/ops.cuh: Side): the forks are joined back
    inside every call.  Replays must reproduce the eagerly computed loss and gradients (fp32 atomics in the weight
    gradients make them equal only up to summation order) and must follow new inputs."""
    import ctypes as C
    from stgcn_b200 import models, _lib as L
    from stgcn_b200.graph import GraphedStep
    dev = cuda_device
    gso = load_gso(dataset, "cheb" if kind == "cheb_graph_conv" else "gcn")
    n = gso.shape[0]
    blocks = [[1], [64, 16, 64], [64, 16, 64], [128, 128], [1]]
    args = SimpleNamespace(Kt=3, Ks=3, act_func="glu", graph_conv_type=kind, gso=gso.to(dev), enable_bias=True,
                           droprate=0.0, n_his=12)
    cls = models.STGCNChebGraphConv if kind == "cheb_graph_conv" else models.STGCNGraphConv
    model = cls(args, blocks, n).to(dev)
    model.load_state_dict(O.init_params(blocks=blocks, kt=3, ks=3, n_his=12, n_vertex=n, kind=kind, seed=5))
    model.train()
    gen = torch.Generator().manual_seed(11)
    xs = [torch.randn(B, 1, 12, n, generator=gen).to(dev) for _ in range(2)]
    ys = [torch.randn(B, n, generator=gen).to(dev) for _ in range(2)]

    def eager(x, y):
        model.zero_grad(set_to_none=True)
        pred = model(x).reshape(B, -1).float()
        dpred = torch.empty_like(pred)
        loss = torch.zeros(1, device=dev)
        L.check(L.lib().stgcn_mse_fwd_bwd(pred.data_ptr(), y.data_ptr(), pred.numel(), C.c_float(1.0), loss.data_ptr(),
                                          dpred.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        pred.backward(dpred)
        torch.cuda.synchronize()
        return loss.item(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}

    ref = [eager(x, y) for x, y in zip(xs, ys)]
    step = GraphedStep(model, (B, 1, 12, n), (B, n), device=dev, warmup=2)
    for rep in range(2):                       # second pass: replays are repeatable
        for (x, y), (loss_ref, grads_ref) in zip(zip(xs, ys), ref):
            loss = step(x, y)
            torch.cuda.synchronize()
            assert abs(loss.item() - loss_ref) <= 1e-5 * abs(loss_ref) + 1e-7, (rep, loss.item(), loss_ref)
            got = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
            assert set(got) == set(grads_ref)
            for k, g_ref in grads_ref.items():
                assert rel_l2(got[k].cpu(), g_ref.cpu()) <= 1e-4, (rep, k, rel_l2(got[k].cpu(), g_ref.cpu()))


def _pems_model(dev, droprate, B, seed=5):
    from stgcn_b200.synthetic import build_model
    gso = load_gso("pemsd7m", "cheb")
    blocks = [[1], [64, 16, 64], [64, 16, 64], [128, 128], [1]]
    model = build_model(gso, "cheb_graph_conv", 3, blocks, dev, droprate=droprate, seed=seed)
    model.train()
    gen = torch.Generator().manual_seed(11)
    n = gso.shape[0]
    return model, torch.randn(B, 1, 12, n, generator=gen).to(dev), torch.randn(B, n, generator=gen).to(dev), n


def test_graphed_step_dropout_masks_change_per_replay(cuda_device):
    """A captured step bakes the by-value dropout seeds in; the device-side step counter (stgcn_set_dropout_step) must
    give every replay fresh masks, and the forward and backward of one replay the SAME mask."""
    from stgcn_b200.graph import GraphedStep
    dev = cuda_device
    B = 1          # 8 (b, t) groups: ~2^-8 of the (n, c) positions are dropped in every group, which the last check needs
    model, x, y, n = _pems_model(dev, 0.5, B)
    # the mask of the last dropout (output block, after fc1+ReLU) decides which fc2 inputs are zero: observe it through
    # a forward hook on the first ST block's output instead (LayerNorm output has no exact zeros without dropout)
    seen = []
    h = model.st_blocks[0].register_forward_hook(lambda m, i, o: seen.append(o))
    step = GraphedStep(model, (B, 1, 12, n), (B, n), device=dev, warmup=2)
    h.remove()
    static_out = seen[-1]                     # the tensor the captured graph writes block 0's output into
    masks, losses, grads = [], [], []
    for rep in range(3):
        loss = step(x, y)
        torch.cuda.synchronize()
        masks.append((static_out != 0).clone())
        losses.append(loss.item())
        grads.append(model.st_blocks[0].tc2_ln.weight.grad.clone())
    for m in masks:
        assert abs(m.float().mean().item() - 0.5) < 0.02
    assert not torch.equal(masks[0], masks[1]) and not torch.equal(masks[1], masks[2])
    assert len({round(l, 6) for l in losses}) == 3
    # forward/backward agreement: dropped positions of block 0's output receive no gradient through the LayerNorm
    # affine, so d(ln.weight) equals the sum over kept positions only -- recompute it from an eager pass with the same
    # counter value is not possible from outside; instead check the necessary condition that a (n, c) column that was
    # dropped in every (b, t) of a replay has an exactly zero weight gradient
    for m, g in zip(masks, grads):
        kept_any = m.permute(0, 2, 3, 1).reshape(-1, m.shape[3], m.shape[1]).any(0)      # (N, C)
        assert torch.count_nonzero(g[~kept_any]) == 0
    step.close()


def test_graphed_step_regrads_after_zero_grad(cuda_device):
    """optimizer.zero_grad(set_to_none=True) between replays must not disconnect p.grad from the graph's buffers."""
    from stgcn_b200.graph import GraphedStep
    dev = cuda_device
    B = 4
    model, x, y, n = _pems_model(dev, 0.0, B)
    step = GraphedStep(model, (B, 1, 12, n), (B, n), device=dev, warmup=2)
    step(x, y)
    torch.cuda.synchronize()
    ref = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    assert all(p.grad is None for p in model.parameters())
    step(x, y)
    torch.cuda.synchronize()
    for k, p in model.named_parameters():
        if k in ref:
            assert p.grad is not None and rel_l2(p.grad.cpu(), ref[k].cpu()) <= 1e-4, k
    step.close()


@pytest.mark.parametrize("N,B,kt,kind", [(228, 5, 3, "cheb_graph_conv"), (41, 3, 3, "cheb_graph_conv"), (207, 2, 2, "graph_conv"),
                                          (325, 3, 3, "cheb_graph_conv")])
def test_bf16_first_block_fused_backward(N, B, kt, kind, cuda_device):
    """Block 0 of the default architecture without a data gradient (the model input needs none): the align conv's data
    gradient, the GLU backward (z recomputed from x) and the first conv's weight gradient run as ONE tcgen05 kernel
    (csrc/umma_fb0.cuh).  Checked against (a) the unfused path -- same block with x.requires_grad -- and (b) the fp64
    oracle; ragged last row tile (rows not a multiple of 128), Kt = 2 and 3."""
    from stgcn_b200 import layers, _lib as L
    dev = cuda_device
    gen = torch.Generator().manual_seed(N + B)
    torch.manual_seed(N + B)
    gso = O.synthetic_gso(N, seed=N)
    T = 12 if kt == 3 else 8
    blk = layers.STConvBlock(kt, 3, N, 1, [64, 16, 64], "glu", kind, gso.to(dev), True, 0.0).to(dev)
    blk.train()
    x = torch.randn(B, 1, T, N, generator=gen)
    dy = torch.randn(B, 64, T - 2 * (kt - 1), N, generator=gen)

    def run(requires_grad):
        blk.zero_grad(set_to_none=True)
        xg = x.to(dev).requires_grad_(requires_grad)
        L.profile_begin()
        blk(xg).backward(dy.to(dev).bfloat16())
        prof = L.profile_end()
        return {k: p.grad.detach().float().cpu().clone() for k, p in blk.named_parameters() if p.grad is not None}, prof

    fused, prof_f = run(False)
    plain, prof_p = run(True)
    assert any("umma_fb0_kernel" in k for k in prof_f), sorted(prof_f)
    assert not any("umma_fb0_kernel" in k for k in prof_p)
    assert not any("lowrank_expand" in k or "smallc1_gate_wgrad" in k for k in prof_f)
    assert set(fused) == set(plain)
    for k in fused:
        tol = 3e-2 if k.startswith("tmp_conv1.causal_conv") else 1e-6        # everything else runs the same kernels
        assert rel_l2(fused[k], plain[k]) <= tol, (k, rel_l2(fused[k], plain[k]))
    # fp64 oracle of the block
    p64 = {"b." + k: v.detach().double().cpu().requires_grad_(True) for k, v in blk.state_dict().items()}
    y64 = O.st_conv_block(x.double(), p64, "b.", gso.double(), kt, [64, 16, 64], "glu", kind)
    y64.backward(dy.double())
    for k in ("tmp_conv1.causal_conv.weight", "tmp_conv1.causal_conv.bias"):
        e_fused, e_plain = rel_l2(fused[k], p64["b." + k].grad), rel_l2(plain[k], p64["b." + k].grad)
        assert e_fused < GRAD_TOL and e_fused < 1.5 * e_plain + 1e-2, (k, e_fused, e_plain)


@pytest.mark.parametrize("N,B,T,kind", [(228, 5, 12, "cheb_graph_conv"), (41, 3, 12, "cheb_graph_conv"), (207, 2, 8, "graph_conv"),
                                         (325, 3, 12, "cheb_graph_conv"), (228, 150, 7, "cheb_graph_conv"), (100, 2, 5, "cheb_graph_conv")])
def test_bf16_second_conv_fused_backward(N, B, T, kind, cuda_device):
    """Blocks of the default architecture (16 -> 64 GLU channels, Kt = 3, no dropout): the LayerNorm backward, the GLU
    backward and the data / weight / bias gradients of the second temporal conv run as ONE tcgen05 kernel
    (csrc/umma_fb2.cuh) behind ln_bwd_sums_pg_kernel.  Checked against the fp64 oracle of the block on every parameter
    gradient and on dx: ragged vertex tiles (N = 228, 325, 41), 1..3 tiles, more samples than CTAs per tile (B = 150: several
    items per CTA, the accumulator rings wrap), T2 from 3 to 8.  With dropout active in training the block must take the
    unfused path (the mask is applied there) and still agree with itself run to run."""
    from stgcn_b200 import layers, _lib as L
    dev = cuda_device
    gen = torch.Generator().manual_seed(N + B + T)
    torch.manual_seed(N + B + T)
    gso = O.synthetic_gso(N, seed=N)
    blk = layers.STConvBlock(3, 3, N, 64, [64, 16, 64], "glu", kind, gso.to(dev), True, 0.0).to(dev)
    blk.train()
    x = torch.randn(B, 64, T, N, generator=gen)
    dy = torch.randn(B, 64, T - 4, N, generator=gen)
    xg = x.to(dev).requires_grad_(True)
    L.profile_begin()
    blk(xg).backward(dy.to(dev).bfloat16())
    prof = L.profile_end()
    assert any("umma_fb2_kernel" in k for k in prof), sorted(prof)
    assert any("ln_bwd_sums_pg_kernel" in k for k in prof)
    assert not any("ln_gate_bwd_kernel" in k for k in prof)
    got = {k: p.grad.detach().float().cpu() for k, p in blk.named_parameters() if p.grad is not None}
    got["dx"] = xg.grad.detach().float().cpu()
    p64 = {"b." + k: v.detach().double().cpu().requires_grad_(True) for k, v in blk.state_dict().items()}
    x64 = x.double().requires_grad_(True)
    O.st_conv_block(x64, p64, "b.", gso.double(), 3, [64, 16, 64], "glu", kind).backward(dy.double())
    ref = {k: p64["b." + k].grad for k in got if k != "dx"}
    ref["dx"] = x64.grad
    for k in got:
        assert ref[k] is not None, k
        assert rel_l2(got[k], ref[k].float()) < GRAD_TOL, (k, rel_l2(got[k], ref[k].float()))
    # the tensors the fused kernel produces itself, tighter: second conv weight / bias, LayerNorm weight / bias
    for k in ("tmp_conv2.causal_conv.weight", "tmp_conv2.causal_conv.bias", "tc2_ln.weight", "tc2_ln.bias"):
        assert rel_l2(got[k], ref[k].float()) < 3e-2, (k, rel_l2(got[k], ref[k].float()))
    # dropout in training: unfused path
    blk_d = layers.STConvBlock(3, 3, N, 64, [64, 16, 64], "glu", kind, gso.to(dev), True, 0.5).to(dev)
    blk_d.train()
    L.profile_begin()
    blk_d(x.to(dev).requires_grad_(True)).backward(dy.to(dev).bfloat16())
    prof_d = L.profile_end()
    assert not any("umma_fb2_kernel" in k for k in prof_d), sorted(prof_d)
