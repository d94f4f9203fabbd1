"""GPU parity of the callers either side of the path: fused flat-buffer AdamW / Lion against the oracle
and against torch.optim on the same gradients; device-side window construction bit exact against data_transform."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import stgcn_oracle as O
from oracle import train_oracle as T

pytestmark = pytest.mark.gpu


def _tiny_model(dev, seed=0):
    from stgcn_b200.synthetic import build_model
    n = 23
    blocks = [[1], [16, 8, 16], [16, 8, 16], [32, 32], [1]]
    gso = O.synthetic_gso(n, seed=2)
    model = build_model(gso, "cheb_graph_conv", 3, blocks, dev, seed=seed)
    model.train()
    gen = torch.Generator().manual_seed(1)
    return model, torch.randn(6, 1, 12, n, generator=gen).to(dev), torch.randn(6, n, generator=gen).to(dev)


@pytest.mark.parametrize("which", ["adamw", "lion"])
def test_flat_optimizer_matches_torch_per_tensor(which, cuda_device):
    """Five training steps of the same model twice: torch's per-tensor optimizer (AdamW) / the oracle (Lion) vs ONE fused
    launch on the flat buffer.  Same gradients by construction (same kernels, same inputs), so the parameters must agree
    to fp32 rounding; dead parameters stay untouched."""
    import stgcn_b200
    from stgcn_b200.optim import FlatAdamW, FlatLion
    dev = cuda_device
    stgcn_b200.set_precision("fp32")
    ma, x, y = _tiny_model(dev)
    mb, _, _ = _tiny_model(dev)
    mb.load_state_dict(ma.state_dict())
    before = {k: v.detach().clone() for k, v in ma.named_parameters()}

    def backward(m):
        m.zero_grad(set_to_none=True)
        torch.nn.functional.mse_loss(m(x).view(x.shape[0], -1), y).backward()

    backward(mb)
    opt_b = FlatAdamW(mb, lr=2e-3, weight_decay=0.05) if which == "adamw" else FlatLion(mb, lr=2e-3, weight_decay=0.05)
    if which == "adamw":
        opt_a = torch.optim.AdamW(ma.parameters(), lr=2e-3, weight_decay=0.05)
    lion_state = {}
    for it in range(5):
        backward(ma)
        if it > 0:
            backward(mb)                       # (step 0's backward already ran: it bound the flat buffer)
        if which == "adamw":
            opt_a.step()
        else:
            with torch.no_grad():
                for k, p in ma.named_parameters():
                    if p.grad is None:
                        continue
                    m = lion_state.get(k, np.zeros(p.numel(), np.float32))
                    pn, m = T.lion_step(p.detach().cpu().numpy().reshape(-1), p.grad.cpu().numpy().reshape(-1), m,
                                        lr=2e-3, betas=(0.9, 0.99), weight_decay=0.05)
                    lion_state[k] = m
                    p.copy_(torch.from_numpy(pn).view_as(p))
        opt_b.step()
        torch.cuda.synchronize()
        pa = dict(ma.named_parameters())
        for k, p in mb.named_parameters():
            if which == "adamw":
                assert rel_l2(p.detach().cpu(), pa[k].detach().cpu()) < 2e-6, (it, k)
            else:      # sign updates: a sign flip moves an element by 2 lr; allow a handful from fp32 reassociation
                d = (p.detach().cpu() - pa[k].detach().cpu()).abs()
                assert (d > 1e-6).float().mean().item() < 2e-3, (it, k)
    live = set(opt_b.reducer.names)
    for k, p in mb.named_parameters():
        if k not in live:
            assert torch.equal(p.detach(), before[k]), k          # dead align convs: untouched
        else:
            assert p.data_ptr() == opt_b.flat_params.data_ptr() + 4 * opt_b.reducer.offsets[opt_b.reducer.names.index(k)]


def test_adamw_kernel_matches_oracle(cuda_device):
    import ctypes as C
    from stgcn_b200 import _lib as L
    dev = cuda_device
    n = 100003                                   # not a multiple of 4: exercises the scalar tail
    g = torch.Generator().manual_seed(0)
    p = torch.randn(n, generator=g)
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    pd, md, vd = p.to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    steps = torch.zeros(1, dtype=torch.int64, device=dev)
    pn = p.numpy().copy()
    for t in range(1, 4):
        gr = torch.randn(n, generator=g) * 0.3
        pn, m, v = T.adamw_step(pn, gr.numpy(), m, v, t, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-3)
        gd = gr.to(dev)
        L.check(L.lib().stgcn_adamw_step(pd.data_ptr(), gd.data_ptr(), md.data_ptr(), vd.data_ptr(), n, C.c_float(1e-3),
                                         C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8), C.c_float(1e-3), C.c_float(1.0),
                                         1, steps.data_ptr(), None, torch.cuda.current_stream().cuda_stream))
        steps.add_(1)
        assert np.allclose(pd.cpu().numpy(), pn, rtol=3e-6, atol=1e-7), t
    assert np.allclose(md.cpu().numpy(), m, rtol=1e-5, atol=1e-7) and np.allclose(vd.cpu().numpy(), v, rtol=1e-4, atol=1e-9)


def test_device_windows_bit_exact(cuda_device):
    from stgcn_b200.data import DeviceWindows
    dev = cuda_device
    z = np.load(os.path.join(GOLDEN, "train_windows.npz"))
    series = torch.from_numpy(z["data"].astype(np.float32)).to(dev)
    n_his, n_pred = int(z["n_his"]), int(z["n_pred"])
    xr, yr = T.data_transform(z["data"].astype(np.float32), n_his, n_pred)
    assert np.array_equal(xr, z["x"]) and np.array_equal(yr, z["y"])          # float32 series: same bits as the golden
    w = DeviceWindows(series, n_his, n_pred)
    assert len(w) == xr.shape[0]
    x, y = w.batch(0)                                                          # the whole split, sequential order
    assert torch.equal(x.cpu(), torch.from_numpy(xr)) and torch.equal(y.cpu(), torch.from_numpy(yr))
    x, y = w.batch(start=7, size=16)
    assert torch.equal(x.cpu(), torch.from_numpy(xr[7:23])) and torch.equal(y.cpu(), torch.from_numpy(yr[7:23]))
    idx = torch.tensor([45, 0, 3, 3, 17], dtype=torch.int64)
    x, y = w.batch(starts=idx)
    assert torch.equal(x.cpu(), torch.from_numpy(xr[idx.numpy()])) and torch.equal(y.cpu(), torch.from_numpy(yr[idx.numpy()]))
    x, y = w.batch(start=40, size=100)                                         # ragged tail like the last DataLoader batch
    assert x.shape[0] == len(w) - 40 and torch.equal(x.cpu(), torch.from_numpy(xr[40:]))


def test_graphed_step_with_fused_optimizer_trains(cuda_device):
    """forward + loss + backward + fused AdamW in ONE captured graph: the loss on a fixed batch goes down over replays
    and the step counter / bias correction advance inside the graph."""
    import stgcn_b200
    from stgcn_b200.graph import GraphedStep
    from stgcn_b200.optim import FlatAdamW
    dev = cuda_device
    stgcn_b200.set_precision("bf16")
    try:
        model, x, y = _tiny_model(dev, seed=3)
        torch.nn.functional.mse_loss(model(x).view(x.shape[0], -1).float(), y).backward()
        opt = FlatAdamW(model, lr=5e-3, weight_decay=0.0)
        step = GraphedStep(model, tuple(x.shape), tuple(y.shape), device=dev, warmup=2, post_backward=opt.step)
        n0 = int(opt.steps_dev.item())
        losses = [step(x, y).item() for _ in range(30)]
        assert int(opt.steps_dev.item()) == n0 + 30
        assert losses[-1] < 0.7 * losses[0], losses[::6]
        step.close()
    finally:
        stgcn_b200.set_precision("fp32")


@pytest.mark.parametrize("name,types", [("pemsd7m", ("sym_norm_lap", "sym_renorm_adj")),
                                         ("rand", ("sym_norm_adj", "sym_renorm_adj", "sym_norm_lap", "sym_renorm_lap",
                                                   "rw_norm_adj", "rw_renorm_lap"))])
def test_device_gso_matches_reference(name, types, cuda_device):
    """Operator preprocessing on the device vs the reference-derived golden operators (sym_* types)
    and the oracle (rw_* types, which crash in the reference under the installed scipy): calc_gso to fp32 rounding,
    the Chebyshev rescale to the accuracy of the power iteration's lambda_max (<= 1e-5 relative)."""
    from stgcn_b200 import gso as G
    dev = cuda_device
    z = np.load(os.path.join(GOLDEN, "train_gso.npz"))
    adj = torch.from_numpy(z[f"adj_{name}"]).to(dev)
    for t in types:
        ref = T.calc_gso_dense(z[f"adj_{name}"].astype(np.float64), t)
        if f"{name}_{t}" in z.files:
            assert np.allclose(ref, z[f"{name}_{t}"], rtol=1e-6, atol=1e-7)
        got = G.calc_gso(adj, t)
        assert rel_l2(got.cpu(), torch.from_numpy(ref)) < 2e-6, t
        cheb_ref, lam = T.calc_chebynet_gso_dense(ref)
        cheb, eig = G.calc_chebynet_gso(got, return_eigval=True)
        assert abs(eig[0].item() - lam) <= 1e-5 * lam, (t, eig.tolist(), lam)
        assert rel_l2(cheb.cpu(), torch.from_numpy(cheb_ref)) < 3e-5, t
        one = G.build_operator(adj, t, True)
        assert rel_l2(one.cpu(), torch.from_numpy(cheb_ref)) < 3e-5, t
    with pytest.raises(ValueError):
        G.calc_gso(adj, "sym_lap")
