# SM-cycle timeline of the tap kernel's roles around tiles 8..10 of CTA 0 (build with -DSTGCN_TIMELINE):
#   tools/build_variants.sh tl="-DSTGCN_TIMELINE" && STGCN_B200_LIB=$PWD/build/variants/tl.so python tools/tap_cycles_probe.py
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, stgcn_b200
from stgcn_b200 import layers, _lib as L
stgcn_b200.set_precision("bf16")
dev = torch.device('cuda')
buf = torch.zeros(128, dtype=torch.int64, device=dev)
LAB = {32: 'mma t8 top', 33: 'mma t9 top', 34: 'mma t10 top', 35: 'mma t11 top', 36: 'mma t8 tempty ok', 37: 'mma t8 issued',
       38: 'mma t8 tfull committed', 39: 'mma t8 end',
       40: 'prod g24 top', 41: 'prod g24 empty ok', 42: 'prod g24 issued', 43: 'prod g25 top', 44: 'prod g25 empty ok',
       45: 'prod g25 issued', 46: 'prod g26 top', 47: 'prod g26 empty ok', 48: 'prod g26 issued',
       49: 'epi2 t8 top', 50: 'epi2 t8 sempty ok', 51: 'epi2 t8 tfull ok', 52: 'epi2 t8 cols done', 53: 'epi2 t8 tempty arrived',
       54: 'epi2 t8 sfull arrived', 55: 'epi2 t9 top', 56: 'epi2 t9 sempty ok', 57: 'epi2 t9 tfull ok', 58: 'epi2 t9 cols done',
       59: 'epi2 t9 tempty arrived', 60: 'epi2 t9 sfull arrived', 61: 'epi2 t10 top', 62: 'epi2 t10 sempty ok',
       63: 'epi2 t10 tfull ok', 64: 'epi2 t10 cols done', 65: 'epi2 t10 tempty arrived', 66: 'epi2 t10 sfull arrived',
       67: 'epiLast t8 top', 68: 'epiLast t9 top',
       70: 'store t8 top', 71: 'store t8 sfull ok', 72: 'store t8 committed', 73: 'store t8 released',
       80: 'mma t8 tap0 ready', 81: 'mma t8 tap0 issued', 82: 'mma t8 tap1 ready', 83: 'mma t8 tap1 issued', 84: 'mma t8 tap2 ready',
       85: 'mma t8 tap2 issued', 86: 'mma t8 tap3 ready', 87: 'mma t8 tap3 issued', 88: 'mma t8 residual issued',
       90: 'cp0 g24 top', 91: 'cp0 g24 empty ok', 92: 'cp0 g24 issued', 93: 'cp0 g24 prev published',
       94: 'cp0 g28 top', 95: 'cp0 g28 empty ok', 96: 'cp0 g28 issued', 97: 'cp0 g28 prev published', 98: 'cp0 g32 top', 99: 'cp0 g32 empty ok',
       100: 'cp0 g32 issued', 101: 'cp0 g32 prev published',
       74: 'store t9 top', 75: 'store t9 sfull ok', 76: 'store t9 committed', 77: 'store t9 released'}
def run(name, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    L.check(L.lib().stgcn_debug_timeline(buf.data_ptr()))
    buf.zero_(); fn(); torch.cuda.synchronize()
    L.check(L.lib().stgcn_debug_timeline(None))
    t = buf.cpu().tolist()
    t0 = t[32]
    print(name, '(SM cycles relative to the issuer\'s tile-8 loop top; 1965 cycles = 1 us)')
    ev = sorted((t[k] - t0, v) for k, v in LAB.items() if t[k] and abs(t[k] - t0) < 10**7)
    for c, v in ev: print(f"   {c:8d}  {v}")
B, N = 256, 228
# the block-level path: q-only GLU state, bias and residual on the tensor pipe (the stamps left in the buffer are those of
# the block's LAST tap launch of the forward = its second temporal conv, 16 -> 64)
gso = torch.eye(N, device=dev)
for (cin, T) in [(64, 8)]:
    blk = layers.STConvBlock(3, 3, N, cin, [64, 16, 64], 'glu', 'cheb_graph_conv', gso, True, 0.0).to(dev)
    x = torch.randn(B, cin, T, N, device=dev)
    with torch.no_grad():
        run(f"STConvBlock fwd c_in={cin} T={T}: tc2 16->64", lambda: blk(x))
tc = layers.TemporalConvLayer(3, 64, 64, N, 'glu').to(dev)      # stand-alone layer: explicit-residual (AUX) variant
x = torch.randn(B, 64, 8, N, device=dev)
run("tconv 64->64 T=8 stand-alone (aux epilogue)", lambda: tc(x))
