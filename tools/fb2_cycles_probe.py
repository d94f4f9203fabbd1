# SM-cycle timeline of umma_fb2_kernel's roles around tiles 4..6 of CTA 0's first item (build with -DSTGCN_TIMELINE)
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, stgcn_b200
from stgcn_b200 import layers, _lib as L
stgcn_b200.set_precision("bf16")
dev = torch.device('cuda')
buf = torch.zeros(256, dtype=torch.int64, device=dev)
LAB = {}
for k in range(3):
    for e, n in enumerate(['top', 'dZ ready', 'H2 ready', 'D3 free', 'dgrad issued', 'wgrad issued']):
        LAB[k * 8 + e] = f'mma t{4 + k} {n}'
    for e, n in enumerate(['E1 top', 'E1 math done', 'E1 dz slot ok', 'E1 stored+arrived']):
        LAB[24 + k * 8 + e] = f'epi2 t{4 + k} {n}'
    LAB[48 + k] = f'epi17 t{4 + k} E1 top'
for k in range(4):
    LAB[52 + 2 * k] = f'E2 tau{3 + k} begin'; LAB[53 + 2 * k] = f'E2 tau{3 + k} read+freed'
B, N = 256, 228
gso = torch.eye(N, device=dev)
blk = layers.STConvBlock(3, 3, N, 64, [64, 16, 64], 'glu', 'cheb_graph_conv', gso, True, 0.0).to(dev)
x = torch.randn(B, 64, 12, N, device=dev, requires_grad=True)
def step():
    y = blk(x); y.backward(torch.ones_like(y))
for _ in range(2): step()
torch.cuda.synchronize()
L.profile_begin(); step(); torch.cuda.synchronize(); prof = L.profile_end()
for k, (n, ms) in sorted(prof.items()):
    if 'fb2' in k or 'sums' in k: print(f"{k:50s} {ms / n * 1000:8.1f} us")
L.check(L.lib().stgcn_debug_timeline(buf.data_ptr()))
buf.zero_(); step(); torch.cuda.synchronize()
L.check(L.lib().stgcn_debug_timeline(None))
t = buf.cpu().tolist()[160:]; t0 = t[0]
raw = buf.cpu().tolist()

print("fb2 (SM cycles relative to the issuer's tile-4 top; 1965 cycles = 1 us)")
for c, v in sorted((t[k] - t0, v) for k, v in LAB.items() if t[k] and abs(t[k] - t0) < 10**7): print(f"   {c:8d}  {v}")
