"""Diagnostics for the fused graph-convolution kernels (csrc/umma_cheb.cuh): calls the C ABI directly and compares
every Chebyshev plane, the output and the input gradient with a torch fp32 evaluation of the same bf16 operands."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stgcn_b200 import _lib as L


def rel(a, b):
    return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))


def run(N, B, T, Ks, kind, relu, residual, seed=0):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    Cc = 16
    a = torch.randn(N, N, generator=g)
    gso = (a / torch.linalg.matrix_norm(a, ord=2)).float().to(dev)
    x = torch.randn(B, T, N, Cc, generator=g).to(dev).bfloat16()
    ntap = Ks if kind == 0 else 1
    w = (torch.randn(ntap, Cc, Cc, generator=g) * 0.3).to(dev)
    b = torch.randn(Cc, generator=g).to(dev)
    lib = L.lib()
    desc = L.GconvDesc(B, T, N, Cc, Cc, Ks, kind, relu, residual, L.PREC["bf16"])
    sv, ws = C.c_size_t(), C.c_size_t()
    L.check(lib.stgcn_gconv_sizes(C.byref(desc), C.byref(sv), C.byref(ws)))
    saved = torch.zeros(sv.value, dtype=torch.uint8, device=dev)
    wsb = torch.zeros(max(ws.value, 256), dtype=torch.uint8, device=dev)
    y = torch.zeros(B, T, N, Cc, dtype=torch.bfloat16, device=dev)
    params = L.GconvParams(None, None, w.data_ptr(), b.data_ptr(), gso.data_ptr())
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.stgcn_gconv_fwd(C.byref(desc), x.data_ptr(), C.byref(params), y.data_ptr(), saved.data_ptr(),
                                wsb.data_ptr(), wsb.numel(), st))
    torch.cuda.synchronize()
    depth = Ks if kind == 0 else 2
    stack = saved.view(torch.bfloat16)[: depth * B * T * N * Cc].view(depth, B, T, N, Cc).float()
    Lb = gso.bfloat16().float()
    xs = [x.float()]
    rb = lambda t: t.bfloat16().float()
    for k in range(1, depth):
        nxt = torch.einsum("hi,btic->bthc", Lb, xs[k - 1])
        if k >= 2:
            nxt = 2 * nxt - xs[k - 2]
        xs.append(rb(nxt))
    wb = rb(w)
    if kind == 0:
        out = sum(torch.einsum("btnc,co->btno", xs[k], wb[k]) for k in range(Ks))
    else:
        out = torch.einsum("btnc,co->btno", xs[1], wb[0])
    out = out + b
    if residual:
        out = out + xs[0]
    if relu:
        out = torch.relu(out)
    msg = [f"N={N} B={B} T={T} Ks={Ks} kind={kind} relu={relu} res={residual}:"]
    for k in range(1, depth):
        msg.append(f"x{k} {rel(stack[k], xs[k]):.2e}")
    msg.append(f"y {rel(y.float(), out):.2e}")
    # backward
    dy = torch.randn(B, T, N, Cc, generator=g).to(dev).bfloat16()
    dx = torch.zeros_like(x)
    gw = torch.zeros_like(w); gb = torch.zeros_like(b)
    grads = L.GconvGrads(None, None, gw.data_ptr(), gb.data_ptr())
    L.check(lib.stgcn_gconv_bwd(C.byref(desc), x.data_ptr(), saved.data_ptr(), dy.data_ptr(), C.byref(params),
                                C.byref(grads), dx.data_ptr(), wsb.data_ptr(), wsb.numel(), st))
    torch.cuda.synchronize()
    # reference adjoint on the kernel's own y (mask) and bf16 operands
    dg = dy.float() * (y.float() > 0) if relu else dy.float()
    D = [None] * depth
    if kind == 0:
        for k in range(Ks):
            D[k] = torch.einsum("btno,co->btnc", dg, wb[k])
    else:
        D[0] = torch.zeros_like(dg); D[1] = torch.einsum("btno,co->btnc", dg, wb[0])
    for k in range(depth - 1, 0, -1):
        alpha = 2.0 if k >= 2 else 1.0
        D[k - 1] = D[k - 1] + alpha * torch.einsum("ih,btic->bthc", Lb, rb(D[k]))
        if k >= 2:
            D[k - 2] = D[k - 2] - rb(D[k])
    dx_ref = D[0] + (dg if residual else 0)
    msg.append(f"dx {rel(dx.float(), dx_ref):.2e}")
    gw_ref = torch.stack([torch.einsum("btnc,btno->co", xs[k if kind == 0 else 1], dg) for k in range(ntap)])
    msg.append(f"gw {rel(gw, gw_ref):.2e} gb {rel(gb, dg.sum((0, 1, 2))):.2e}")
    print(" ".join(msg), flush=True)


if __name__ == "__main__":
    torch.manual_seed(0)
    run(41, 2, 3, 3, 0, 1, 1)
    run(41, 2, 3, 2, 0, 0, 0)
    run(41, 2, 3, 3, 1, 1, 1)
    run(130, 2, 3, 3, 0, 1, 1)
    run(228, 3, 5, 3, 0, 1, 1)
    run(228, 3, 5, 5, 0, 0, 1)
    run(207, 2, 7, 3, 1, 1, 1)
    run(228, 64, 10, 3, 0, 1, 1)
