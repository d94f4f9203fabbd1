"""In-kernel timeline of the fused graph-convolution kernels (CTA 0, first 3 items of each slot), from the
globaltimer stamps of csrc/umma_cheb.cuh.  Diagnostics only: prints microseconds relative to the kernel start."""
# NOTE: the in-kernel globaltimer stamps are compiled in only with -DSTGCN_TIMELINE:
#   tools/build_variants.sh tl="-DSTGCN_TIMELINE" && STGCN_B200_LIB=$PWD/build/variants/tl.so python <this script>

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, stgcn_b200
from stgcn_b200 import layers, _lib as L

stgcn_b200.set_precision("bf16")
dev = torch.device("cuda")
buf = torch.zeros(256, dtype=torch.int64, device=dev)
B, N, T, depth = 256, 228, int(os.environ.get("T", 10)), 3


def show(name):
    t = buf.cpu().tolist()
    t0 = t[150]
    us = lambda v: round((v - t0) / 1e3, 2) if v else None
    print(f"== {name}: operator resident at {us(t[151])} us")
    for slot in range(2):
        for it in range(3):
            b = slot * 72 + it * 24
            print(f" slot {slot} item {it}: fill {us(t[b])}->{us(t[b+1])} | mma issued",
                  [us(t[b + 2 + j]) for j in range(depth)], "| epi",
                  [(us(t[b + 8 + 2 * j]), us(t[b + 9 + 2 * j])) for j in range(depth)],
                  "| hop waits-done/issued", [(us(t[b + 16 + j]), us(t[b + 20 + j])) for j in range(depth - 1)],
                  "| epi0 detail: tile0 tmem", us(t[b + 14]), "tile1 tmem", us(t[b + 19]), "stores issued", us(t[b + 15]), "fence done", us(t[b + 23]))


gl = layers.GraphConvLayer("cheb_graph_conv", 16, 16, 3, torch.randn(N, N, device=dev) / 30, True).to(dev)
x = torch.randn(B, 16, T, N, device=dev).bfloat16().requires_grad_(True)
for _ in range(3):
    y = gl(x, _relu=1); y.backward(torch.ones_like(y))
torch.cuda.synchronize()
L.check(L.lib().stgcn_debug_timeline(buf.data_ptr()))
buf.zero_(); y = gl(x, _relu=1); torch.cuda.synchronize()
show("forward")
buf.zero_(); y.backward(torch.ones_like(y)); torch.cuda.synchronize()
show("backward")
L.check(L.lib().stgcn_debug_timeline(None))
