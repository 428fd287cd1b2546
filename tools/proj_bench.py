"""Micro-benchmark of the short-K projection products (development aid): fhvae_proj_planes_batch (TMA, planes) vs
fhvae_gemm_batch (gemm_tc.cu, fp32 operands), graph-replayed, at the shapes of the config-1 step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from pytorch_scalablefhvae_b200 import _lib
from pytorch_scalablefhvae_b200._lib import GemmProblem, ProjProblem
from pytorch_scalablefhvae_b200.plan import ptr, gemm_nt

def graph_us(fn, reps=10):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): fn(s.cuda_stream)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(reps): fn(torch.cuda.current_stream().cuda_stream)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best

for (M, N, K) in [(5120, 1024, 80), (5120, 160, 256)]:
    A, W, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") * 0.1, torch.randn(N, device="cuda")
    C = torch.zeros(M, N, device="cuda")
    pl = lambda t: torch.stack([t.to(torch.bfloat16).reshape(-1), (t - t.to(torch.bfloat16).float()).to(torch.bfloat16).reshape(-1)]).contiguous()
    pa, pw = pl(A), pl(W)
    for mode in (1, 2):
        pp = (ProjProblem * 1)(ProjProblem(pa.data_ptr(), pw.data_ptr(), ptr(C), ptr(b), M, N, K, 0, K, M * K, K, N * K, N))
        gp = (GemmProblem * 1)(gemm_nt(ptr(A), K, ptr(W), K, ptr(C), N, M, N, K, bias=ptr(b)))
        t1 = graph_us(lambda st: _lib.check(_lib.fn("fhvae_proj_planes_batch")(pp, 1, mode, st)))
        t2 = graph_us(lambda st: _lib.check(_lib.fn("fhvae_gemm_batch")(gp, 1, mode, st)))
        print(f"{M}x{N}x{K} mode {mode}: proj_planes {t1:.1f} us   gemm_tc {t2:.1f} us   (output {M * N * 4 / 1e6:.1f} MB)")
