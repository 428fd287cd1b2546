"""Shared helpers for the GPU parity tests (tests may import oracle/, the package may not)."""
import numpy as np
import torch

from pytorch_scalablefhvae_b200 import _lib
from pytorch_scalablefhvae_b200._lib import ColsumProblem, GemmProblem
from pytorch_scalablefhvae_b200.plan import ptr

FP32_RTOL = 1e-4      # north_star: fp32 lower bound / KL terms / gradients within 1e-4 relative
BF16_RTOL = 2e-2      # north_star: bf16 input-GEMM mode, stated separately


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    st = _lib.fn(name)(*args, stream())
    _lib.check(st, name)


def relerr(a: torch.Tensor, b: torch.Tensor) -> float:
    """max-norm relative error of a against the reference b (SURVEY.md Appendix D calibration)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = float(b.abs().max())
    if denom == 0.0:
        return float(a.abs().max())
    return float((a - b).abs().max()) / denom


def assert_close(a, b, rtol=FP32_RTOL, what=""):
    assert tuple(a.shape) == tuple(b.shape), f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    e = relerr(a, b)
    assert e <= rtol, f"{what}: max-norm relative error {e:.3e} > {rtol:.1e}"


def gemm(A, B, C, M, N, K, sa, sb, ldc, bias=None, beta=0.0, relu=0, mode=0):
    p = GemmProblem(ptr(A), ptr(B), ptr(C), ptr(bias) if bias is not None else None, M, N, K, relu,
                    sa[0], sa[1], sb[0], sb[1], ldc, beta, 0)
    arr = (GemmProblem * 1)(p)
    call("fhvae_gemm_batch", arr, 1, mode)


def synth_batch(B, T, F, N, seed=1234):
    """SURVEY.md §8d synthetic inputs: x ~ N(0,1); utterance lengths U[200,1600] -> nsegs; idx ∝ nsegs."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, F, generator=g)
    rng = np.random.default_rng(7)
    lens = rng.integers(200, 1601, size=N)
    nsegs_u = (lens - 20) // 8 + 1
    p = nsegs_u / nsegs_u.sum()
    idx = torch.from_numpy(np.random.default_rng(seed).choice(N, size=B, p=p)).long()
    nsegs = torch.from_numpy(nsegs_u)[idx].long()
    return x, idx, nsegs
