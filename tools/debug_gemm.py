#!/usr/bin/env python
"""Debug aid: every GEMM of the GwEpilogue chain against torch float64, twice (bitwise repeatability = race check)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from step_b200 import ops  # noqa: E402

DEV = "cuda:0"


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp_min(1e-12))


def check(name, fn, ref):
    o1 = fn()
    o2 = fn()
    print(f"{name:40s} err {rel(o1, ref):.2e}  repeat-equal {bool(torch.equal(o1, o2))}", flush=True)
    return o1


for M in (69, 414, 883, 6624):
    print("M =", M)
    g = torch.Generator().manual_seed(M)
    h, skip = torch.randn(M, 96, generator=g), torch.randn(M, 256, generator=g)
    shapes = [(512, 96), (512,), (256, 512), (256,), (512, 256), (512,), (12, 512), (12,)]
    w1, b1, w2, b2, we1, be1, we2, be2 = [torch.randn(*s, generator=g) / math.sqrt(s[-1] if len(s) > 1 else 16.0) for s in shapes]
    dout = torch.randn(M, 12, generator=g)
    D = lambda t: t.double()
    h1r = torch.relu(D(h) @ D(w1).t() + D(b1))
    hsr = torch.relu(h1r @ D(w2).t() + D(b2))
    x2r = torch.relu(hsr + D(skip))
    e1r = torch.relu(x2r @ D(we1).t() + D(be1))
    de1r = (D(dout) @ D(we2)) * (e1r > 0)
    dx2r = (de1r @ D(we1)) * (x2r > 0)
    dz2r = dx2r * (hsr > 0)
    dh1r = (dz2r @ D(w2)) * (h1r > 0)
    G = lambda t: t.float().to(DEV)
    hd, skd, w1d, b1d, w2d, b2d, we1d, be1d, we2d, be2d, doutd = map(G, (h, skip, w1, b1, w2, b2, we1, be1, we2, be2, dout))
    h1 = check("h1 = relu(h w1^T + b1)", lambda: ops.gemm(hd, w1d, bias=b1d, epilogue=ops.GE_RELU), h1r)
    hs = torch.empty(M, 256, device=DEV)
    x2 = check("x2 = relu(relu(h1 w2^T+b2)+skip)", lambda: ops.gemm(G(h1r), w2d, bias=b2d, epilogue=ops.GE_RELU_ADD_RELU, aux=skd, aux_out=hs), x2r)
    print(f"{'   hs (aux_out)':40s} err {rel(hs, hsr):.2e}")
    check("e1 = relu(x2 we1^T + be1)", lambda: ops.gemm(G(x2r), we1d, bias=be1d, epilogue=ops.GE_RELU), e1r)
    check("de1 = (dout we2) * [e1>0]", lambda: ops.gemm(doutd, we2d, transB=True, epilogue=ops.GE_MASK, aux=G(e1r)), de1r)
    check("dx2 = (de1 we1) * [x2>0]", lambda: ops.gemm(G(de1r), we1d, transB=True, epilogue=ops.GE_MASK, aux=G(x2r)), dx2r)
    check("dx2 nomask = de1 we1", lambda: ops.gemm(G(de1r), we1d, transB=True), de1r @ D(we1))
    check("dh1 = (dz2 w2) * [h1>0]", lambda: ops.gemm(G(dz2r), w2d, transB=True, epilogue=ops.GE_MASK, aux=G(h1r)), dh1r)
    for ks in (1, 2, 4):
        check(f"dw1 = dh1^T h  (ksplit {ks})", lambda: ops.gemm(G(dh1r), hd, transA=True, transB=True, ksplit=ks), dh1r.t() @ D(h))
        check(f"dwe1 = de1^T x2 (ksplit {ks})", lambda: ops.gemm(G(de1r), G(x2r), transA=True, transB=True, ksplit=ks), de1r.t() @ x2r)
        check(f"dwe2 = dout^T e1 (ksplit {ks})", lambda: ops.gemm(doutd, G(e1r), transA=True, transB=True, ksplit=ks), D(dout).t() @ e1r)
    check("colsum(dout)", lambda: ops.colsum(doutd), D(dout).sum(0))

print("==== GwEpilogue autograd Function, all gradients ====")
for M in (69, 414, 883):
    for rep in range(2):
        g = torch.Generator().manual_seed(M + 1000)
        h, skip = torch.randn(M, 96, generator=g), torch.randn(M, 256, generator=g)
        shapes = [(512, 96), (512,), (256, 512), (256,), (512, 256), (512,), (12, 512), (12,)]
        ps = [torch.randn(*s, generator=g) / math.sqrt(s[-1] if len(s) > 1 else 16.0) for s in shapes]
        dout = torch.randn(M, 12, generator=g)
        L = [t.double().requires_grad_(True) for t in ps]
        sk = skip.double().requires_grad_(True)
        hs = torch.relu(torch.relu(h.double() @ L[0].t() + L[1]) @ L[2].t() + L[3])
        ref = torch.relu(torch.relu(sk + hs) @ L[4].t() + L[5]) @ L[6].t() + L[7]
        ref.backward(dout.double())
        mp = [t.to(DEV).requires_grad_(True) for t in ps]
        ms = skip.to(DEV).requires_grad_(True)
        out = ops.GwEpilogue.apply(h.to(DEV), ms, *mp)
        out.backward(dout.to(DEV))
        names = ["w1", "b1", "w2", "b2", "we1", "be1", "we2", "be2"]
        line = f"M={M} rep={rep} out {rel(out.detach(), ref.detach()):.1e} dskip {rel(ms.grad, sk.grad):.1e} " + \
            " ".join(f"{n} {rel(m.grad, l.grad):.1e}" for n, m, l in zip(names, mp, L))
        print(line, flush=True)
