"""tcgen05.mma cost table on one SM (stgcn_umma_microbench): cycles per instruction to issue and to complete, by shape,
operand layout and number of independent accumulator chains.  Output is committed under profiles/."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stgcn_b200 import _lib as L

NONE, S128, S64, S32 = 0, 2, 4, 6
KMAJ128 = (S128, 16, 1024, 32)          # swizzle, lbo, sbo, k advance: K-major, 64-element K blocks
KMAJ32 = (S32, 16, 256, 0)              # K-major, one 16-element K block
MN128_A = (S128, 16384, 1024, 2048)     # MN-major, 64-wide chunks 16 KB apart (128-row K tile)
MN128_B = (S128, 16384, 1024, 2048)
MN32_G = (S32, 7680, 256, 512)          # MN-major, one 16-wide atom per group, groups 7680 B apart (cheb buffers)
MN32_1 = (S32, 0, 256, 512)

cfgs = []
def add(name, M, N, a_mn, b_mn, a_tmem, A, B, chains, style=0, warps=1):
    for ch in chains:
        if 128 + warps * ch * N <= 512:
            cfgs.append((f"{name} N={N} chains={ch} style={style} warps={warps}",
                         [M, N, a_mn, b_mn, a_tmem, *A, *B, 64, ch, N, style | (warps << 4)]))

for N in (16, 64, 256):
    add("SS  A K-major sw128 | B K-major sw128 ", 128, N, 0, 0, 0, KMAJ128, KMAJ128, (1,))
for style in (1, 2):
    for N in (16, 64, 256):
        add("SS  A K-major sw128 | B K-major sw128 ", 128, N, 0, 0, 0, KMAJ128, KMAJ128, (1,), style=style)
for warps in (2, 4):
    for N in (16, 64):
        add("SS  A K-major sw128 | B K-major sw128 ", 128, N, 0, 0, 0, KMAJ128, KMAJ128, (1,), style=0, warps=warps)
        add("SS  A K-major sw128 | B K-major sw128 ", 128, N, 0, 0, 0, KMAJ128, KMAJ128, (1,), style=2, warps=warps)
add("TS  A tmem         | B MN-major sw32  ", 128, 32, 0, 1, 1, KMAJ128, MN32_G, (1, 2))
add("SS  A MN-major sw128| B MN-major sw32  ", 128, 16, 1, 1, 0, MN128_A, MN32_1, (1,), style=2)
add("SS  A MN-major sw128| B MN-major sw128 ", 128, 64, 1, 1, 0, MN128_A, MN128_B, (1,), style=2)
add("SS  A K-major sw128 | B MN-major sw32  ", 128, 64, 0, 1, 0, KMAJ128, MN32_G, (1,), style=2)

lib = L.lib()
dev = torch.device("cuda:0")
out = torch.zeros(4, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
print(f"{'configuration':62s} {'issue cyc/mma':>14s} {'total cyc/mma':>14s}   (64 instructions, M=128 unless noted, K=16, bf16)")
for name, c in cfgs:
    arr = (C.c_int32 * 17)(*c)
    L.check(lib.stgcn_umma_microbench(arr, out.data_ptr(), st))
    torch.cuda.synchronize()
    o = out.cpu().tolist()
    print(f"{name:62s} {o[0] / o[2]:14.1f} {o[1] / o[2]:14.1f}")
