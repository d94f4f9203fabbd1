"""Generates tests/golden/train_lion.npz and train_windows.npz from the UNMODIFIED reference (script/opt.py Lion,
script/dataloader.py data_transform) on seeded inputs.  Run in the build container (where /root/reference is mounted):
    python tests/golden/make_train_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from script import opt as ref_opt, dataloader as ref_dl      # noqa: E402


def main():
    g = torch.Generator().manual_seed(0)
    # ---- Lion: 4 steps on one tensor, fixed gradients
    p0 = torch.randn(513, generator=g)
    grads = [torch.randn(513, generator=g) * (0.1 if i % 2 else 1.0) for i in range(4)]
    grads[2][:7] = 0.0                               # exercises sign(0) = 0 once the momentum is seeded with zeros ... 
    p = torch.nn.Parameter(p0.clone())
    o = ref_opt.Lion([p], lr=3e-3, betas=(0.9, 0.99), weight_decay=0.02)
    traj = []
    for gr in grads:
        p.grad = gr.clone()
        o.step()
        traj.append(p.detach().clone().numpy())
    np.savez(os.path.join(HERE, "train_lion.npz"), p0=p0.numpy(), grads=np.stack([x.numpy() for x in grads]),
             traj=np.stack(traj), exp_avg=o.state[p]["exp_avg"].numpy(), lr=3e-3, b1=0.9, b2=0.99, wd=0.02)
    # ---- windows
    rng = np.random.default_rng(1)
    data = rng.standard_normal((61, 13))
    x, y = ref_dl.data_transform(data, 12, 3, "cpu")
    np.savez(os.path.join(HERE, "train_windows.npz"), data=data, x=x.numpy(), y=y.numpy(), n_his=12, n_pred=3)
    print("written", x.shape, y.shape)


if __name__ == "__main__":
    main()


def make_gso_golden():
    """Adjacency inputs for the device-side operator preprocessing with the operators the UNMODIFIED
    reference derives from them (utility.calc_gso / calc_chebynet_gso, sym_* types: the rw_* types crash in the reference
    under the installed scipy).  PeMSD7-M's real adjacency (the dense form of data/pemsd7-m/adj.npz) and a
    small random directed weighted graph with an isolated vertex."""
    import scipy.sparse as sp
    from script import utility
    os.chdir("/root/reference")
    adj, n = ref_dl.load_adj("pemsd7-m")
    dense = np.asarray(adj.todense(), dtype=np.float32)
    out = {"adj_pemsd7m": dense}
    for t in ("sym_norm_lap", "sym_renorm_adj"):
        g = utility.calc_gso(sp.csc_matrix(dense.astype(np.float64)), t)
        out[f"pemsd7m_{t}"] = np.asarray(g.todense(), dtype=np.float32)
        out[f"pemsd7m_{t}_cheb"] = np.asarray(utility.calc_chebynet_gso(g).todense(), dtype=np.float32)
    rng = np.random.default_rng(3)
    a = rng.random((40, 40)) * (rng.random((40, 40)) < 0.2)
    np.fill_diagonal(a, 0.0)
    a[7, :] = 0.0
    a[:, 7] = 0.0                                   # isolated vertex: degree 0 -> 1/0 must become 0
    out["adj_rand"] = a.astype(np.float32)
    for t in ("sym_norm_adj", "sym_renorm_adj", "sym_norm_lap", "sym_renorm_lap"):
        g = utility.calc_gso(sp.csc_matrix(a.astype(np.float32).astype(np.float64)), t)
        out[f"rand_{t}"] = np.asarray(g.todense(), dtype=np.float32)
        out[f"rand_{t}_cheb"] = np.asarray(utility.calc_chebynet_gso(g).todense(), dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "train_gso.npz"), **out)
    print("written train_gso.npz", {k: v.shape for k, v in out.items() if k.startswith("adj")})


if __name__ == "__main__" and "--gso" in sys.argv:
    make_gso_golden()
