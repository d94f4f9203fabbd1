"""Driver for profiling the graph-convolution layer alone (ncu target): a few forward+backward passes at the
PeMSD7-M block-0 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, stgcn_b200
from stgcn_b200 import layers

stgcn_b200.set_precision("bf16")
dev = torch.device("cuda")
B, N, T = 256, 228, int(os.environ.get("T", 10))
c_in = int(os.environ.get("CIN", 16))
gl = layers.GraphConvLayer("cheb_graph_conv", c_in, 16, 3, torch.randn(N, N, device=dev) / 30, True).to(dev)
x = torch.randn(B, c_in, T, N, device=dev).bfloat16().requires_grad_(True)
for _ in range(int(os.environ.get("ITERS", 4))):
    y = gl(x, _relu=1)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("done")
