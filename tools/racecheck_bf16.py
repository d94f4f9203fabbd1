# bf16-mode-only racecheck target: one tiny model step (forward + backward) through the tcgen05 kernels
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, stgcn_b200
from types import SimpleNamespace
from stgcn_b200 import models, synthetic
stgcn_b200.set_precision("bf16")
dev = torch.device("cuda:0"); torch.manual_seed(0)
n, B = 37, 4
blocks = [[1], [64, 16, 64], [64, 16, 64], [128, 128], [1]]
gso = synthetic.synthetic_operator(n, seed=3)
model = synthetic.build_model(gso, "cheb_graph_conv", 3, blocks, dev)
model.train()
x = torch.randn(B, 1, 12, n, device=dev); y = torch.randn(B, n, device=dev)
loss = torch.nn.functional.mse_loss(model(x).view(B, -1).float(), y); loss.backward(); torch.cuda.synchronize()
print("bf16 step done, loss", float(loss))
