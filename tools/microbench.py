"""Per-kernel micro-timings on one B200 (CUDA events, L2 flushed between reps): LSTM recurrence
kernels vs T, and every GEMM shape of the config-1 step.  Development aid; numbers quoted in
profiles/ come from here."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_scalablefhvae_b200 as P
from pytorch_scalablefhvae_b200 import _lib
from pytorch_scalablefhvae_b200._lib import GemmProblem
from pytorch_scalablefhvae_b200.plan import ptr

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


COLD = bool(int(os.environ.get("COLD", "0")))


def timeit(f, reps=5, inner=20):
    """Median device time of one call: `inner` back-to-back launches are captured in a CUDA graph so
    that host launch latency (ctypes + cudaLaunch, ~10-20 us) is not what gets measured."""
    f(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            f()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if COLD:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / inner)
    ts.sort()
    return ts[len(ts) // 2] * 1e3   # us


def st():
    return torch.cuda.current_stream().cuda_stream


def lstm(mode, T, B=256, H=256):
    z = lambda *s: torch.randn(*s, device=dev) * 0.3
    Pp, Q, W = z(T, B, 4 * H), z(B, 4 * H), z(4 * H, H) / 16
    h, c, a = z(T, B, H), z(T, B, H), z(T, B, 4 * H)
    dg, dgs, dhr, dc, dh = z(T, B, 4 * H), z(B, 4 * H), z(16, B, H), z(B, H), z(T, B, H)
    fw = lambda: _lib.check(_lib.fn("fhvae_lstm_fwd")(ptr(Pp), ptr(Q), ptr(W), ptr(h), ptr(c), ptr(a), ptr(dhr), T, B, H, mode, st()))
    bw = lambda: _lib.check(_lib.fn("fhvae_lstm_bwd")(ptr(dh), None, ptr(W), ptr(c), ptr(a), ptr(dg), ptr(dgs), ptr(dhr), ptr(dc), T, B, H, mode, st()))
    return timeit(fw), timeit(bw)


def wave(mode, T, L, B=256, H=256):
    z = lambda *s: torch.randn(*s, device=dev) * 0.3
    Pp, Q = z(T, B, 4 * H), z(B, 4 * H)
    W0, Wi1, W1, b1 = z(4 * H, H) / 16, z(4 * H, H) / 16, z(4 * H, H) / 16, z(4 * H)
    o = [[z(T, B, H), z(T, B, H), z(T, B, 4 * H)] for _ in range(2)]
    n = _lib.fn("fhvae_lstm_wave_xchg_bytes")(T, B, H, L)
    xchg = torch.zeros(n // 4, device=dev)
    l1 = [ptr(Wi1), ptr(b1), ptr(W1), ptr(o[1][0]), ptr(o[1][1]), ptr(o[1][2])] if L == 2 else [None] * 6
    fw = lambda: _lib.check(_lib.fn("fhvae_lstm_wave_fwd")(ptr(Pp), ptr(Q), ptr(W0), ptr(o[0][0]), ptr(o[0][1]), ptr(o[0][2]),
                                                          *l1, ptr(xchg), T, B, H, L, mode, st()))
    nb = _lib.fn("fhvae_lstm_wave_bwd_xchg_bytes")(T, B, H, L)
    xb = torch.zeros(nb // 4, device=dev)
    dh, dhl, dhl0 = z(T, B, H), z(B, H), z(B, H)
    dg = [z(T, B, 4 * H), z(T, B, 4 * H)]
    dgs = [z(B, 4 * H), z(B, 4 * H)]
    top = L - 1
    bot = [ptr(Wi1), ptr(W0), ptr(o[0][1]), ptr(o[0][2]), ptr(dg[0]), ptr(dgs[0])] if L == 2 else [None] * 6
    Wt = W1 if L == 2 else W0
    bw = lambda: _lib.check(_lib.fn("fhvae_lstm_wave_bwd")(ptr(dh), ptr(dhl), ptr(dhl0) if L == 2 else None, ptr(Wt), ptr(o[top][1]),
                                                          ptr(o[top][2]), ptr(dg[top]), ptr(dgs[top]), *bot, ptr(xb), T, B, H, L, mode, st()))
    return timeit(fw), timeit(bw)


def gemm(mode, M, N, K, kind):
    A = torch.randn(M, K, device=dev); Bm = torch.randn(N, K, device=dev); C = torch.zeros(M, N, device=dev)
    if kind == "nt":
        p = GemmProblem(ptr(A), ptr(Bm), ptr(C), None, M, N, K, 0, K, 1, 1, K, N, 0.0, 0)
    elif kind == "nn":
        Bm = torch.randn(K, N, device=dev)
        p = GemmProblem(ptr(A), ptr(Bm), ptr(C), None, M, N, K, 0, K, 1, N, 1, N, 0.0, 0)
    else:  # tn: C[M,N] = G[K,M]^T X[K,N]
        A = torch.randn(K, M, device=dev); Bm = torch.randn(K, N, device=dev)
        p = GemmProblem(ptr(A), ptr(Bm), ptr(C), None, M, N, K, 0, 1, M, N, 1, N, 0.0, 0)
    arr = (GemmProblem * 1)(p)
    us = timeit(lambda: _lib.check(_lib.fn("fhvae_gemm_batch")(arr, 1, mode, st())))
    return us, 2.0 * M * N * K / us / 1e6   # TFLOP/s


if __name__ == "__main__":
    out = {}
    for mode in (1, 2):
        for T in (1, 2, 5, 20):
            out[f"lstm mode{mode} T{T} fwd/bwd us"] = lstm(mode, T)
    for mode in (1, 2):
        for T, L in ((1, 1), (20, 1), (1, 2), (5, 2), (20, 2)):
            out[f"wave mode{mode} T{T} L{L} fwd/bwd us"] = wave(mode, T, L)
    if os.environ.get("ONLY_LSTM"):
        for k, v in out.items():
            print(k, [round(x, 2) for x in v])
        sys.exit(0)
    shapes = [("nt", 5120, 1024, 80), ("nt", 5120, 1024, 256), ("nt", 5120, 160, 256), ("nt", 256, 1024, 64),
              ("nt", 256, 64, 256), ("nn", 5120, 256, 1024), ("nn", 5120, 256, 160), ("nn", 256, 64, 1024),
              ("tn", 1024, 256, 4864), ("tn", 1024, 256, 5120), ("tn", 1024, 80, 5120), ("tn", 160, 256, 5120),
              ("tn", 1024, 64, 256), ("tn", 64, 256, 256)]
    for mode in (1, 2):
        for kind, M, N, K in shapes:
            out[f"gemm mode{mode} {kind} {M}x{N}x{K} us,TF"] = gemm(mode, M, N, K, kind)
    for k, v in out.items():
        print(k, [round(x, 2) for x in v])
