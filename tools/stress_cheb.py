"""Repeatability stress for the fused graph-convolution kernels: the same problem many times, with the shared
workspace dirtied in between; y / dx must be bit-identical, parameter gradients (fp32 atomics) within 1e-5."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stgcn_b200 import _lib as L


def run(N, B, T, Ks, kind, relu, residual, reps=30, seed=0):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    Cc = 16
    a = torch.randn(N, N, generator=g)
    gso = (a / torch.linalg.matrix_norm(a, ord=2)).float().to(dev)
    x = torch.randn(B, T, N, Cc, generator=g).to(dev).bfloat16()
    ntap = Ks if kind == 0 else 1
    w = (torch.randn(ntap, Cc, Cc, generator=g) * 0.3).to(dev)
    b = torch.randn(Cc, generator=g).to(dev)
    dy = torch.randn(B, T, N, Cc, generator=g).to(dev).bfloat16()
    lib = L.lib()
    desc = L.GconvDesc(B, T, N, Cc, Cc, Ks, kind, relu, residual, L.PREC["bf16"])
    sv, ws = C.c_size_t(), C.c_size_t()
    L.check(lib.stgcn_gconv_sizes(C.byref(desc), C.byref(sv), C.byref(ws)))
    st = torch.cuda.current_stream().cuda_stream
    first = None
    bad = 0
    for rep in range(reps):
        saved = torch.empty(sv.value, dtype=torch.uint8, device=dev)
        wsb = torch.empty(max(ws.value, 256), dtype=torch.uint8, device=dev)
        if rep % 2:
            saved.random_(0, 255); wsb.random_(0, 255)       # garbage (incl. NaN patterns) in every scratch byte
        y = torch.empty(B, T, N, Cc, dtype=torch.bfloat16, device=dev)
        params = L.GconvParams(None, None, w.data_ptr(), b.data_ptr(), gso.data_ptr())
        L.check(lib.stgcn_gconv_fwd(C.byref(desc), x.data_ptr(), C.byref(params), y.data_ptr(), saved.data_ptr(),
                                    wsb.data_ptr(), wsb.numel(), st))
        dx = torch.empty_like(x)
        gw = torch.zeros_like(w); gb = torch.zeros_like(b)
        grads = L.GconvGrads(None, None, gw.data_ptr(), gb.data_ptr())
        if rep % 2:
            wsb.random_(0, 255)
        L.check(lib.stgcn_gconv_bwd(C.byref(desc), x.data_ptr(), saved.data_ptr(), dy.data_ptr(), C.byref(params),
                                    C.byref(grads), dx.data_ptr(), wsb.data_ptr(), wsb.numel(), st))
        torch.cuda.synchronize()
        cur = (y.clone(), dx.clone(), gw.clone(), gb.clone())
        if first is None:
            first = cur
            continue
        ey = (cur[0].float() - first[0].float()).abs().max().item()
        ex = (cur[1].float() - first[1].float()).abs().max().item()
        ew = ((cur[2] - first[2]).abs().max() / first[2].abs().max()).item()
        eb = ((cur[3] - first[3]).abs().max() / first[3].abs().max()).item()
        if ey > 0 or ex > 0 or ew > 1e-4 or eb > 1e-4 or not torch.isfinite(cur[1].float()).all():
            bad += 1
            print(f"  rep {rep}: y {ey:.3e} dx {ex:.3e} gw {ew:.3e} gb {eb:.3e}", flush=True)
    print(f"N={N} B={B} T={T} Ks={Ks} kind={kind} relu={relu}: {bad} / {reps - 1} repetitions differ", flush=True)


if __name__ == "__main__":
    run(41, 2, 7, 5, 0, 1, 1)
    run(41, 2, 7, 3, 0, 1, 1)
    run(228, 3, 5, 3, 0, 1, 1)
    run(228, 16, 10, 3, 0, 1, 1)
    run(207, 2, 9, 3, 1, 1, 1)
    run(130, 4, 3, 2, 0, 0, 1)
