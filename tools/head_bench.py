#!/usr/bin/env python
"""Isolated timing of the fused latent-head kernels (heads.cu) vs the number of CTAs / sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_scalablefhvae_b200 import _lib
from pytorch_scalablefhvae_b200.plan import ptr

dev = "cuda"
st = lambda: torch.cuda.current_stream().cuda_stream


def timeit(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


H, L, Z = 256, 2, 32
for B in (2, 32, 128, 256):
    for Kq, NQ, withq in ((32, 1024, 1), (64, 1024, 1), (32, 1024, 0)):
        hs = [torch.randn(B, H, device=dev) for _ in range(L)]
        W, b, eps = torch.randn(2 * Z, L * H, device=dev), torch.randn(2 * Z, device=dev), torch.randn(B, Z, device=dev)
        zcat = torch.randn(B, 64, device=dev)
        Wq = torch.randn(NQ, 112, device=dev)
        head, Q = torch.zeros(B, 2 * Z, device=dev), torch.zeros(B, NQ, device=dev)
        f = lambda: _lib.check(_lib.fn("fhvae_head_fwd")(ptr(hs[0]), ptr(hs[1]), H, L, H, ptr(W), ptr(b), ptr(head), Z, ptr(eps),
                                                         ptr(zcat), 64, 32 if Kq == 32 else 0, ptr(Wq) if withq else None, 112, None,
                                                         32 if Kq == 32 else 0, Kq, ptr(Q) if withq else None, NQ, B, st()))
        t = timeit(f)
        dg = torch.randn(B, 1024, device=dev)
        dz = torch.zeros(B, 64, device=dev)
        dhead = torch.zeros(B, 2 * Z, device=dev)
        dh = [torch.zeros(B, H, device=dev) for _ in range(L)]
        g = lambda: _lib.check(_lib.fn("fhvae_head_bwd")(ptr(dg) if withq else None, 1024, ptr(Wq), 112, Kq, ptr(dz), 64, 0, 0, ptr(head),
                                                         ptr(eps), Z, 0, ptr(dhead), 1, ptr(W), L, H, ptr(dh[0]), ptr(dh[1]), B, st()))
        t2 = timeit(g)
        print(f"B={B:4d} Kq={Kq} withq={withq}: head_fwd {t:7.1f} us   head_bwd {t2:7.1f} us")
