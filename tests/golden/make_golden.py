#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

It imports hazdzz/STGCN's own ``model/models.py`` + ``model/layers.py`` and
``script/utility.py`` + ``script/dataloader.py``, runs seeded forward+backward steps
(the body of main.py:165-168 without the optimizer) and writes, per case, one
``case_<name>.npz`` holding: the config (json), inputs ``x``/``y``, the graph shift
operator ``gso``, every state_dict entry (``p:<key>``), the model output, the output of
the first ST block, the loss, ``dx`` and every parameter gradient (``g:<key>``; keys
whose gradient is None in the reference -- the dead align convs -- are absent).
It also writes ``gso_<dataset>_<type>.npy``: the dense fp32 operators the reference
derives from its shipped adjacency files (main.py:97-103), used by the parity tests and
by bench.py as realistic L-hat inputs (the GPU box has no /root/reference).
"""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.path.insert(0, REF)
    os.chdir(REF)   # dataloader uses ./data relative paths (dataloader.py:8)
    from model import models          # noqa
    from script import utility, dataloader   # noqa
    return models, utility, dataloader


def small_gso(n, seed, symmetric=True):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(n, n, generator=g, dtype=torch.float64)
    if symmetric:
        a = (a + a.T) / 2
    a = a / torch.linalg.matrix_norm(a, ord=2)
    return a.float()


CASES = [
    # name, dict(cfg)
    ("tiny_cheb3_glu", dict(n=20, B=3, blocks=[[1], [8, 4, 8], [8, 4, 8], [16, 16], [1]], Kt=3, Ks=3, act="glu", kind="cheb_graph_conv", bias=True, n_his=12)),
    ("tiny_cheb3_gtu", dict(n=20, B=2, blocks=[[1], [8, 4, 8], [8, 4, 8], [16, 16], [1]], Kt=3, Ks=3, act="gtu", kind="cheb_graph_conv", bias=True, n_his=12)),
    ("tiny_cheb2_relu", dict(n=17, B=2, blocks=[[1], [8, 4, 8], [8, 4, 8], [16, 16], [1]], Kt=3, Ks=2, act="relu", kind="cheb_graph_conv", bias=True, n_his=12)),
    ("tiny_cheb5_silu", dict(n=17, B=2, blocks=[[1], [8, 4, 8], [8, 4, 8], [16, 16], [1]], Kt=3, Ks=5, act="silu", kind="cheb_graph_conv", bias=True, n_his=12)),
    ("tiny_cheb1_glu_nobias", dict(n=13, B=2, blocks=[[1], [8, 4, 8], [8, 4, 8], [16, 16], [1]], Kt=3, Ks=1, act="glu", kind="cheb_graph_conv", bias=False, n_his=12)),
    ("tiny_gcn_glu", dict(n=19, B=3, blocks=[[1], [8, 4, 8], [8, 4, 8], [16, 16], [1]], Kt=3, Ks=3, act="glu", kind="graph_conv", bias=True, n_his=12)),
    ("tiny_gcn_glu_nobias_nonsym", dict(n=19, B=2, blocks=[[1], [8, 4, 8], [8, 4, 8], [16, 16], [1]], Kt=3, Ks=3, act="glu", kind="graph_conv", bias=False, n_his=12, nonsym=True)),
    ("tiny_cheb3_glu_nonsym", dict(n=16, B=2, blocks=[[1], [8, 4, 8], [8, 4, 8], [16, 16], [1]], Kt=3, Ks=3, act="glu", kind="cheb_graph_conv", bias=True, n_his=12, nonsym=True)),
    # Kt=2 -> Ko = 12 - 2*2*1 = 8
    ("tiny_kt2_cheb3_glu", dict(n=15, B=2, blocks=[[1], [8, 4, 8], [8, 4, 8], [16, 16], [1]], Kt=2, Ks=3, act="glu", kind="cheb_graph_conv", bias=True, n_his=12)),
    # channel-adapter variety: tmp_conv with c_in>c_out (1x1 conv residual), graph align pad / identity
    ("tiny_align_mix", dict(n=14, B=2, blocks=[[1], [12, 4, 12], [8, 8, 6], [10, 7], [1]], Kt=3, Ks=3, act="glu", kind="cheb_graph_conv", bias=True, n_his=12)),
    ("tiny_align_pad_gc", dict(n=14, B=2, blocks=[[1], [4, 8, 4], [4, 8, 12], [10, 7], [2]], Kt=3, Ks=2, act="gtu", kind="cheb_graph_conv", bias=True, n_his=12)),
    # Ko == 1 path of models.py:36-51 (raw ST-block output).  Ko == 0 is unreachable in the
    # reference: it implies a zero-length time axis and the last temporal conv raises first.
    ("tiny_ko1", dict(n=11, B=2, blocks=[[1], [8, 4, 8], [8, 4, 8], [16, 16], [1]], Kt=3, Ks=3, act="glu", kind="cheb_graph_conv", bias=True, n_his=9)),
    # one batch element, one block
    ("tiny_one_block", dict(n=9, B=1, blocks=[[1], [8, 4, 8], [16, 16], [1]], Kt=3, Ks=3, act="glu", kind="cheb_graph_conv", bias=True, n_his=12)),
    # real graph, real sizes (PeMSD7-M, the BASELINE.json config), tiny batch
    ("pemsd7m_cheb3_glu", dict(dataset="pemsd7-m", B=2, blocks=[[1], [64, 16, 64], [64, 16, 64], [128, 128], [1]], Kt=3, Ks=3, act="glu", kind="cheb_graph_conv", bias=True, n_his=12)),
]


def main():
    models, utility, dataloader = _import_reference()

    # ---- dense graph shift operators from the shipped adjacency files ------------------
    gsos = {}
    for ds in ("pemsd7-m", "metr-la", "pems-bay"):
        adj, n_vertex = dataloader.load_adj(ds)
        lap = utility.calc_chebynet_gso(utility.calc_gso(adj, "sym_norm_lap")).toarray().astype(np.float32)
        ren = utility.calc_gso(adj, "sym_renorm_adj").toarray().astype(np.float32)
        assert lap.shape == (n_vertex, n_vertex)
        tag = ds.replace("-", "")
        np.save(os.path.join(HERE, f"gso_{tag}_cheb.npy"), lap)
        np.save(os.path.join(HERE, f"gso_{tag}_gcn.npy"), ren)
        gsos[ds] = lap
        print(ds, n_vertex, "cheb/gcn operators saved")

    for idx, (name, cfg) in enumerate(CASES):
        torch.manual_seed(1000 + idx)
        if "dataset" in cfg:
            gso = torch.from_numpy(gsos[cfg["dataset"]])
            n = gso.shape[0]
        else:
            n = cfg["n"]
            gso = small_gso(n, 7 + idx, symmetric=not cfg.get("nonsym", False))
        args = SimpleNamespace(Kt=cfg["Kt"], Ks=cfg["Ks"], act_func=cfg["act"], graph_conv_type=cfg["kind"],
                               gso=gso, enable_bias=cfg["bias"], droprate=0.0, n_his=cfg["n_his"])
        cls = models.STGCNChebGraphConv if cfg["kind"] == "cheb_graph_conv" else models.STGCNGraphConv
        model = cls(args, cfg["blocks"], n)
        # default LayerNorm affine is (1, 0): perturb so the affine path is pinned too
        with torch.no_grad():
            for k, v in model.named_parameters():
                if "_ln." in k:
                    v.add_(0.1 * torch.randn_like(v))
        model.train()
        B = cfg["B"]
        x = torch.randn(B, 1, cfg["n_his"], n, requires_grad=True)
        out = model(x)
        y = torch.randn(B, out.numel() // B)
        block0 = model.st_blocks[0](x)
        loss = torch.nn.functional.mse_loss(out.reshape(B, -1), y)   # main.py:166-167
        loss.backward()
        rec = {"cfg": np.array(json.dumps({**cfg, "n": n})),
               "x": x.detach().numpy(), "y": y.numpy(), "gso": gso.numpy(),
               "out": out.detach().contiguous().numpy(), "block0_out": block0.detach().contiguous().numpy(),
               "loss": np.array(loss.item(), dtype=np.float64), "dx": x.grad.numpy()}
        for k, v in model.state_dict().items():
            rec["p:" + k] = v.numpy()
        for k, v in model.named_parameters():
            if v.grad is not None:
                rec["g:" + k] = v.grad.numpy()
        path = os.path.join(HERE, f"case_{name}.npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: out{tuple(out.shape)} loss={loss.item():.6f} -> {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
