# Knock-out probe of the tap kernel: run with STGCN_B200_LIB pointing at a build with -DSTGCN_KO_{STORE,EPI,LOAD,MMA}
# (tools/build_variants.sh) and compare the per-launch times -- which resource bounds the 128-row tile period.
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, stgcn_b200
from stgcn_b200 import layers, _lib as L
stgcn_b200.set_precision("bf16")
dev = torch.device('cuda')
B, N = 256, 228
tag = os.environ.get("KO_TAG", "base")
for (cin, cout, T) in [(16, 64, 10), (64, 64, 8)]:
    lay = layers.TemporalConvLayer(3, cin, cout, N, 'glu').to(dev)
    x = torch.randn(B, cin, T, N, device=dev, requires_grad=True)
    for _ in range(3):
        y = lay(x); y.backward(torch.ones_like(y))
    torch.cuda.synchronize()
    L.profile_begin()
    for _ in range(10):
        y = lay(x); y.backward(torch.ones_like(y))
    torch.cuda.synchronize()
    prof = L.profile_end()
    for k, (n, ms) in sorted(prof.items()):
        if 'umma' in k:
            print(f"{tag:16s} tconv {cin}->{cout} T={T}  {k:60s} {ms / n * 1000:8.1f} us  x{n // 10}")
