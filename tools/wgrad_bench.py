#!/usr/bin/env python
"""Isolated timing of the weight-gradient GEMMs of one LSTM stack: TMA/bf16-plane kernel (gemm_wgrad.cu, with and
without the split pass) against the register-staged fp32 kernel (gemm_tc.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_scalablefhvae_b200 import _lib
from pytorch_scalablefhvae_b200._lib import GemmProblem, SplitProblem, WgradProblem
from pytorch_scalablefhvae_b200.plan import ptr, gemm_tn

dev = "cuda"
st = lambda: torch.cuda.current_stream().cuda_stream
T, B, H, F = 20, 256, 256, 80
TB = T * B


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


dg = [torch.randn(TB, 4 * H, device=dev) for _ in range(2)]
h = [torch.randn(TB, H, device=dev) for _ in range(2)]
x = torch.randn(TB, F, device=dev)
out = [torch.zeros(4 * H, H, device=dev) for _ in range(3)] + [torch.zeros(4 * H, F, device=dev)]
out2 = [torch.zeros_like(o) for o in out]
# fp32 path (as the plan issues it)
probs = [gemm_tn(ptr(dg[1], B * 4 * H), 4 * H, ptr(h[1]), H, ptr(out[0]), H, 4 * H, H, (T - 1) * B),
         gemm_tn(ptr(dg[1]), 4 * H, ptr(h[0]), H, ptr(out[1]), H, 4 * H, H, TB),
         gemm_tn(ptr(dg[0], B * 4 * H), 4 * H, ptr(h[0]), H, ptr(out[2]), H, 4 * H, H, (T - 1) * B),
         gemm_tn(ptr(dg[0]), 4 * H, ptr(x), F, ptr(out[3]), F, 4 * H, F, TB)]
arr = (GemmProblem * 4)(*probs)
t_old = timeit(lambda: _lib.check(_lib.fn("fhvae_gemm_batch")(arr, 4, 1, st())))
# planes
pl = lambda t: torch.zeros(2, t.numel(), dtype=torch.bfloat16, device=dev)
pdg, ph, px = [pl(t) for t in dg], [pl(t) for t in h], pl(x)
sp = [SplitProblem(ptr(t), p.data_ptr(), t.shape[1], t.shape[1], t.numel(), t.shape[0], t.shape[1])
      for t, p in zip(dg + h + [x], pdg + ph + [px])]
sarr = (SplitProblem * len(sp))(*sp)
t_split = timeit(lambda: _lib.check(_lib.fn("fhvae_split_planes_batch")(sarr, len(sp), st())))
e = 2
ND, NH, NX = dg[0].numel(), h[0].numel(), x.numel()
wp = [WgradProblem(pdg[1].data_ptr() + B * 4 * H * e, ph[1].data_ptr(), ptr(out2[0]), 4 * H, H, (T - 1) * B, 0, 4 * H, ND, H, NH, H),
      WgradProblem(pdg[1].data_ptr(), ph[0].data_ptr(), ptr(out2[1]), 4 * H, H, TB, 0, 4 * H, ND, H, NH, H),
      WgradProblem(pdg[0].data_ptr() + B * 4 * H * e, ph[0].data_ptr(), ptr(out2[2]), 4 * H, H, (T - 1) * B, 0, 4 * H, ND, H, NH, H),
      WgradProblem(pdg[0].data_ptr(), px.data_ptr(), ptr(out2[3]), 4 * H, F, TB, 0, 4 * H, ND, F, NX, F)]
warr = (WgradProblem * 4)(*wp)
t_new = timeit(lambda: _lib.check(_lib.fn("fhvae_wgrad_planes_batch")(warr, 4, 1, st())))
err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(out2, out))
flops = 2.0 * 4 * H * TB * (3 * H + F) * 3
print(f"fp32-staged gemm_tc: {t_old:.1f} us | split pass: {t_split:.1f} us | TMA planes: {t_new:.1f} us "
      f"({flops / t_new / 1e6:.0f} TFLOP/s of bf16 MMA work) | max rel diff vs gemm_tc {err:.2e}")
