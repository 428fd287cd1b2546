"""GPU: each C-ABI kernel against a torch fp64/fp32 statement of the same op (sizes the CPU does in
seconds) + exact-index / edge cases (duplicates, first/last row, ragged sizes)."""
import math

import numpy as np
import pytest
import torch

from oracle import fhvae_oracle as O
from pytorch_scalablefhvae_b200 import _lib
from pytorch_scalablefhvae_b200._lib import ColsumProblem, GemmProblem, SplitProblem, WgradProblem
from pytorch_scalablefhvae_b200.plan import ptr
from util import FP32_RTOL, assert_close, call, gemm, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda"


_XCH = {}


def XCH(B, H):
    """exchange scratch (16,B,H) of the cluster LSTM kernels, kept alive for the whole test session"""
    if (B, H) not in _XCH:
        _XCH[B, H] = torch.zeros(16, B, H, device=DEV)
    return _XCH[B, H]


def rnd(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).to(DEV)


# ------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(64, 64, 16), (5120, 1024, 80), (37, 129, 53), (256, 64, 512), (1, 7, 3)])
def test_gemm_nt_bias_relu(M, N, K):
    A, W, b = rnd(M, K, seed=1), rnd(N, K, seed=2), rnd(N, seed=3)
    Cc = torch.zeros(M, N, device=DEV)
    gemm(A, W, Cc, M, N, K, (K, 1), (1, K), N, bias=b, relu=1)
    ref = torch.relu(A.double() @ W.double().t() + b.double())
    assert_close(Cc, ref, 1e-5, "gemm_nt")


def test_gemm_nn_tn_beta_strided():
    M, N, K = 200, 96, 300
    G, W = rnd(M, K, seed=4), rnd(K, N + 8, seed=5)              # W has ld N+8, use cols [4, 4+N)
    C0 = rnd(M, N, seed=6)
    Cc = C0.clone()
    p = GemmProblem(ptr(G), ptr(W, 4), ptr(Cc), None, M, N, K, 0, K, 1, N + 8, 1, N, 1.0, 0)
    call("fhvae_gemm_batch", (GemmProblem * 1)(p), 1, 0)
    assert_close(Cc, C0.double() + G.double() @ W[:, 4:4 + N].double(), 1e-5, "gemm_nn beta")
    # wgrad: dW[N,K2] = dY[R,N]^T X[R,K2]
    R, Nn, K2 = 777, 40, 24
    dY, X = rnd(R, Nn, seed=7), rnd(R, K2, seed=8)
    dW = torch.zeros(Nn, K2, device=DEV)
    gemm(dY, X, dW, Nn, K2, R, (1, Nn), (K2, 1), K2)
    assert_close(dW, dY.double().t() @ X.double(), 1e-5, "gemm_tn")


def test_gemm_grouped_launch():
    shapes = [(100, 64, 32), (5, 200, 77), (300, 16, 16)]
    probs, outs, refs = [], [], []
    for i, (M, N, K) in enumerate(shapes):
        A, W = rnd(M, K, seed=10 + i), rnd(N, K, seed=20 + i)
        Cc = torch.zeros(M, N, device=DEV)
        probs.append(GemmProblem(ptr(A), ptr(W), ptr(Cc), None, M, N, K, 0, K, 1, 1, K, N, 0.0, 0))
        outs.append(Cc); refs.append(A.double() @ W.double().t()); outs.append(A); outs.append(W)
    call("fhvae_gemm_batch", (GemmProblem * 3)(*probs), 3, 0)
    for i in range(3):
        assert_close(outs[3 * i], refs[i], 1e-5, f"group {i}")


TC_TOL = {1: 2e-5, 2: 1e-2}     # bf16x3 split (fp32-parity mode) / single bf16 pass


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (5120, 1024, 80), (5120, 1024, 256), (37, 129, 53),
                                   (256, 64, 512), (300, 160, 16), (256, 1024, 32)])
def test_gemm_tc_nt(M, N, K, mode):
    A, W, b = rnd(M, K, seed=1), rnd(N, K, seed=2), rnd(N, seed=3)
    Cc = torch.full((M, N), 3.0, device=DEV)
    gemm(A, W, Cc, M, N, K, (K, 1), (1, K), N, bias=b, relu=1, mode=mode)
    ref = torch.relu(A.double() @ W.double().t() + b.double())
    assert_close(Cc, ref, TC_TOL[mode], f"gemm_tc nt mode {mode}")


@pytest.mark.parametrize("mode", [1, 2])
def test_gemm_tc_dgrad_wgrad_splitk(mode):
    # dgrad (NN), K = 1024 -> split-K with atomics, beta = 0 and beta = 1
    M, N, K = 5120, 256, 1024
    G, W = rnd(M, K, seed=4), rnd(K, N, seed=5, scale=0.1)
    Cc = torch.full((M, N), 7.0, device=DEV)
    gemm(G, W, Cc, M, N, K, (K, 1), (N, 1), N, mode=mode)
    ref = G.double() @ W.double()
    assert_close(Cc, ref, TC_TOL[mode], "dgrad")
    C1 = torch.ones(M, N, device=DEV)
    gemm(G, W, C1, M, N, K, (K, 1), (N, 1), N, beta=1.0, mode=mode)
    assert_close(C1, ref + 1.0, TC_TOL[mode], "dgrad beta=1")
    # wgrad (TN): dW[1024,256] = dY[4864,1024]^T X[4864,256], into a strided sub-block
    R, Nn, K2 = 4864, 1024, 256
    dY, X = rnd(R, Nn, seed=7, scale=0.1), rnd(R, K2, seed=8)
    dW = torch.full((Nn, K2 + 32), 5.0, device=DEV)
    p = GemmProblem(ptr(dY), ptr(X), ptr(dW, 16), None, Nn, K2, R, 0, 1, Nn, K2, 1, K2 + 32, 0.0, 0)
    call("fhvae_gemm_batch", (GemmProblem * 1)(p), 1, mode)
    assert_close(dW[:, 16:16 + K2], dY.double().t() @ X.double(), TC_TOL[mode], "wgrad")
    assert float((dW[:, :16] - 5.0).abs().max()) == 0.0 and float((dW[:, 16 + K2:] - 5.0).abs().max()) == 0.0


def test_gemm_tc_grouped_mixed():
    shapes = [(100, 64, 32), (5, 200, 77), (300, 16, 8), (640, 256, 1280)]
    probs, keep, refs = [], [], []
    for i, (M, N, K) in enumerate(shapes):
        A, W = rnd(M, K, seed=10 + i), rnd(N, K, seed=20 + i)
        Cc = torch.zeros(M, N, device=DEV)
        probs.append(GemmProblem(ptr(A), ptr(W), ptr(Cc), None, M, N, K, 0, K, 1, 1, K, N, 0.0, 0))
        keep += [A, W, Cc]; refs.append(A.double() @ W.double().t())
    call("fhvae_gemm_batch", (GemmProblem * 4)(*probs), 4, 1)
    for i in range(4):
        assert_close(keep[3 * i + 2], refs[i], 2e-5, f"group {i}")


# ------------------------------------------------------------------------------- TMA weight-gradient GEMM on bf16 planes
def _planes(x):
    """(rows, cols) fp32 -> (2, rows, cols) bf16 planes through the library's split kernel"""
    rows, cols = x.shape
    ps = (rows * cols + 7) // 8 * 8                      # plane stride: multiple of 8 elements
    dst = torch.zeros(2, ps, dtype=torch.bfloat16, device=DEV)
    p = SplitProblem(ptr(x), dst.data_ptr(), x.stride(0), cols, ps, rows, cols)
    call("fhvae_split_planes_batch", (SplitProblem * 1)(p), 1)
    return dst[:, :rows * cols].view(2, rows, cols), dst, ps


def test_split_planes_exact():
    x = rnd(77, 40, seed=1, scale=3.0)
    pl = _planes(x)[0]
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    assert torch.equal(pl[0], hi) and torch.equal(pl[1], lo)
    # strided source (a column slice of a wider matrix)
    w = rnd(33, 64, seed=2)
    pl2 = _planes(w[:, 16:48])[0]
    assert torch.equal(pl2[0], w[:, 16:48].to(torch.bfloat16))


@pytest.mark.parametrize("M,N,K,mode,tol", [(1024, 256, 5120, 1, 2e-5), (1024, 256, 4864, 1, 2e-5), (1024, 80, 5120, 1, 2e-5),
                                           (160, 256, 5120, 1, 2e-5), (64, 256, 256, 1, 2e-5), (72, 24, 100, 1, 2e-5), (5120, 1024, 88, 1, 2e-5),
                                           (304, 520, 333, 1, 2e-5), (1024, 256, 5120, 2, 1e-2)])
def test_wgrad_planes_matches_fp64(M, N, K, mode, tol):
    """C = A^T B from pre-split planes (hi*hi + hi*lo + lo*hi) against fp64, ragged M/N/K and split-K included."""
    A, Bm = rnd(K, M, seed=1), rnd(K, N, seed=2)
    (_, pa, psa), (_, pb, psb) = _planes(A), _planes(Bm)
    ldc = N + 4
    Cc = torch.full((M, ldc), 7.0, device=DEV)
    p = WgradProblem(pa.data_ptr(), pb.data_ptr(), ptr(Cc), M, N, K, 0, M, psa, N, psb, ldc)
    call("fhvae_wgrad_planes_batch", (WgradProblem * 1)(p), 1, mode)
    ref = A.double().t() @ Bm.double()
    assert_close(Cc[:, :N], ref, tol, f"wgrad_planes {M}x{N}x{K}")
    assert torch.all(Cc[:, N:] == 7.0)                   # nothing outside the N columns is touched


def test_wgrad_planes_grouped_and_shifted():
    """One launch, several problems, including the dW_hh form: A rows shifted by one time step against B."""
    T, Bt, H = 6, 64, 256
    dg, h, x = rnd(T * Bt, 4 * H, seed=1), rnd(T * Bt, H, seed=2), rnd(T * Bt, 80, seed=3)
    (_, pdg, ps_dg), (_, ph, ps_h), (_, px, ps_x) = _planes(dg), _planes(h), _planes(x)
    c_hh, c_ih, c_x = (torch.zeros(4 * H, H, device=DEV), torch.zeros(4 * H, H, device=DEV), torch.zeros(4 * H, 112, device=DEV))
    e = 2                                                 # bytes per bf16
    probs = [WgradProblem(pdg.data_ptr() + Bt * 4 * H * e, ph.data_ptr(), ptr(c_hh), 4 * H, H, (T - 1) * Bt, 0, 4 * H, ps_dg, H, ps_h, H),
             WgradProblem(pdg.data_ptr(), ph.data_ptr(), ptr(c_ih), 4 * H, H, T * Bt, 0, 4 * H, ps_dg, H, ps_h, H),
             WgradProblem(pdg.data_ptr(), px.data_ptr(), ptr(c_x), 4 * H, 80, T * Bt, 0, 4 * H, ps_dg, 80, ps_x, 112)]
    call("fhvae_wgrad_planes_batch", (WgradProblem * 3)(*probs), 3, 1)
    assert_close(c_hh, dg[Bt:].double().t() @ h[:-Bt].double(), 2e-5, "dW_hh")
    assert_close(c_ih, dg.double().t() @ h.double(), 2e-5, "dW_ih")
    assert_close(c_x[:, :80], dg.double().t() @ x.double(), 2e-5, "dW_x")
    assert torch.all(c_x[:, 80:] == 0)


# ------------------------------------------------------------------------------- LSTM
def _lstm_ref(P, Q, W, R_all, R_last):
    """fp64 autograd statement of the recurrence used for both fwd and bwd checks."""
    T, B, H4 = P.shape
    H = H4 // 4
    h = torch.zeros(B, H, dtype=torch.float64)
    c = torch.zeros(B, H, dtype=torch.float64)
    hs, cs = [], []
    for t in range(T):
        g = P[t] + Q + h @ W.t()
        i, f, gg, o = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]), torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
        c = f * c + i * gg
        h = o * torch.tanh(c)
        hs.append(h); cs.append(c)
    hs, cs = torch.stack(hs), torch.stack(cs)
    loss = (hs * R_all).sum() + (hs[-1] * R_last).sum()
    return hs, cs, loss


@pytest.mark.parametrize("T,B,H", [(5, 19, 16), (20, 64, 64), (20, 256, 256)])
def test_lstm_fwd_bwd_vs_fp64(T, B, H):
    g = torch.Generator().manual_seed(T * 1000 + B)
    P = (torch.randn(T, B, 4 * H, generator=g) * 0.7).double().requires_grad_(True)
    Q = (torch.randn(B, 4 * H, generator=g) * 0.3).double().requires_grad_(True)
    W = (torch.randn(4 * H, H, generator=g) / math.sqrt(H)).double().requires_grad_(True)
    R_all = torch.randn(T, B, H, generator=g).double()
    R_last = torch.randn(B, H, generator=g).double()
    hs, cs, loss = _lstm_ref(P, Q, W, R_all, R_last)
    loss.backward()
    d = lambda t: t.detach().float().to(DEV).contiguous()
    f = lambda *s: torch.zeros(*s, device=DEV)
    h_all, c_all, acts = f(T, B, H), f(T, B, H), f(T, B, 4 * H)
    Pd, Qd, Wd = d(P), d(Q), d(W)
    call("fhvae_lstm_fwd", ptr(Pd), ptr(Qd), ptr(Wd), ptr(h_all), ptr(c_all), ptr(acts), ptr(XCH(B, H)), T, B, H, 0)
    assert_close(h_all, hs, 2e-5, "h_all")
    assert_close(c_all, cs, 2e-5, "c_all")
    dg, dgsum, dh_rec, dc = f(T, B, 4 * H), f(B, 4 * H), f(16, B, H), f(B, H)
    Ra, Rl = d(R_all), d(R_last)          # keep alive: ptr() of a temporary dangles
    call("fhvae_lstm_bwd", ptr(Ra), ptr(Rl), ptr(Wd), ptr(c_all), ptr(acts), ptr(dg), ptr(dgsum),
         ptr(dh_rec), ptr(dc), T, B, H, 0)
    assert_close(dg, P.grad, 5e-5, "dgates")
    assert_close(dgsum, Q.grad, 5e-5, "dgsum")
    # dW_hh = dg[1:]^T h[:-1] through the library GEMM
    dW = f(4 * H, H)
    gemm(dg[1:], h_all, dW, 4 * H, H, (T - 1) * B, (1, 4 * H), (H, 1), H)
    assert_close(dW, W.grad, 5e-5, "dW_hh")


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("T,B", [(1, 32), (3, 64), (20, 256)])
def test_lstm_cluster_fwd_matches_simt(T, B, mode):
    """Persistent cluster/tcgen05 recurrence (H=256) against the exact fp32 per-step kernels."""
    H = 256
    P = rnd(T, B, 4 * H, seed=1, scale=0.7)
    Q = rnd(B, 4 * H, seed=2, scale=0.3)
    W = rnd(4 * H, H, seed=3, scale=1.0 / 16)
    f = lambda *s: torch.zeros(*s, device=DEV)
    ref = [f(T, B, H), f(T, B, H), f(T, B, 4 * H)]
    out = [f(T, B, H), f(T, B, H), f(T, B, 4 * H)]
    call("fhvae_lstm_fwd", ptr(P), ptr(Q), ptr(W), ptr(ref[0]), ptr(ref[1]), ptr(ref[2]), ptr(XCH(B, H)), T, B, H, 0)
    call("fhvae_lstm_fwd", ptr(P), ptr(Q), ptr(W), ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(XCH(B, H)), T, B, H, mode)
    torch.cuda.synchronize()
    for a, b, n in zip(out, ref, ["h_all", "c_all", "acts"]):
        assert_close(a, b, TC_TOL[mode], f"{n} mode {mode}")


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("T,B,use_all,use_last", [(1, 32, True, True), (4, 64, False, True), (20, 256, True, False),
                                                  (20, 256, True, True)])
def test_lstm_cluster_bwd_matches_simt(T, B, use_all, use_last, mode):
    H = 256
    f = lambda *s: torch.zeros(*s, device=DEV)
    P, Q = rnd(T, B, 4 * H, seed=1, scale=0.7), rnd(B, 4 * H, seed=2, scale=0.3)
    W = rnd(4 * H, H, seed=3, scale=1.0 / 16)
    h_all, c_all, acts = f(T, B, H), f(T, B, H), f(T, B, 4 * H)
    call("fhvae_lstm_fwd", ptr(P), ptr(Q), ptr(W), ptr(h_all), ptr(c_all), ptr(acts), ptr(XCH(B, H)), T, B, H, 0)
    dh_all, dh_last = rnd(T, B, H, seed=4), rnd(B, H, seed=5)
    res = []
    for md in (0, mode):
        dg, dgsum, dh_rec, dc = f(T, B, 4 * H), f(B, 4 * H), f(16, B, H), f(B, H)
        call("fhvae_lstm_bwd", ptr(dh_all) if use_all else None, ptr(dh_last) if use_last else None, ptr(W),
             ptr(c_all), ptr(acts), ptr(dg), ptr(dgsum), ptr(dh_rec), ptr(dc), T, B, H, md)
        torch.cuda.synchronize()
        res.append((dg, dgsum))
    assert_close(res[1][0], res[0][0], TC_TOL[mode], f"dgates mode {mode}")
    assert_close(res[1][1], res[0][1], TC_TOL[mode], f"dgsum mode {mode}")


def _wave_xchg(T, B, H, L):
    n = _lib.fn("fhvae_lstm_wave_xchg_bytes")(T, B, H, L)
    assert n > 0
    return torch.zeros(n // 4, device=DEV)


def _wave_packed(W0, Wi1, W1, H, L, mode):
    nb = _lib.fn("fhvae_lstm_wave_pack_bytes")(H, L, mode)
    ng = H // 32
    assert nb == (2 * ng * L + (2 * ng if L == 2 else 0)) * (2 if mode == 1 else 1) * (H // 128) * 32768
    buf = torch.zeros(nb // 4, device=DEV)
    call("fhvae_lstm_wave_pack", ptr(W0), ptr(Wi1) if Wi1 is not None else None, ptr(W1) if W1 is not None else None,
         ptr(buf), H, L, mode)
    return buf


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("T,B,L,repeat,H", [(1, 32, 1, 1, 256), (3, 64, 2, 2, 256), (20, 256, 1, 2, 256), (20, 256, 2, 3, 256),
                                             (7, 320, 2, 1, 256), (20, 608, 1, 1, 256),
                                             (1, 32, 1, 1, 128), (3, 64, 2, 2, 128), (20, 256, 2, 3, 128), (20, 256, 1, 2, 128),
                                             (5, 608, 2, 1, 128)])
def test_lstm_wave_fwd_matches_simt(T, B, L, repeat, mode, H):
    """Layer-wavefront recurrence (one launch for the stack, LL exchange through L2, layer-1 projection
    in-kernel) against the exact fp32 per-layer kernels + fp32 projection GEMM.  Repeated launches reuse the
    exchange buffer (per-CTA launch counters make the flags unique).  H = 256: groups of 8 CTAs; H = 128 (the
    reference's CLI default width, train_model.py:145-168): groups of 4."""
    P = rnd(T, B, 4 * H, seed=1, scale=0.7)
    Q = rnd(B, 4 * H, seed=2, scale=0.3)
    W0 = rnd(4 * H, H, seed=3, scale=1.0 / 16)
    Wi1, W1, b1 = rnd(4 * H, H, seed=4, scale=1.0 / 16), rnd(4 * H, H, seed=5, scale=1.0 / 16), rnd(4 * H, seed=6, scale=0.2)
    f = lambda *s: torch.zeros(*s, device=DEV)
    ref = [[f(T, B, H), f(T, B, H), f(T, B, 4 * H)] for _ in range(2)]
    call("fhvae_lstm_fwd", ptr(P), ptr(Q), ptr(W0), ptr(ref[0][0]), ptr(ref[0][1]), ptr(ref[0][2]), None, T, B, H, 0)
    if L == 2:
        P1 = f(T, B, 4 * H)
        gemm(ref[0][0], Wi1, P1, T * B, 4 * H, H, (H, 1), (1, H), 4 * H, bias=b1)
        call("fhvae_lstm_fwd", ptr(P1), None, ptr(W1), ptr(ref[1][0]), ptr(ref[1][1]), ptr(ref[1][2]), None, T, B, H, 0)
    xchg = _wave_xchg(T, B, H, L)
    for rep in range(repeat):
        out = [[f(T, B, H), f(T, B, H), f(T, B, 4 * H)] for _ in range(2)]
        l1 = [ptr(Wi1), ptr(b1), ptr(W1), ptr(out[1][0]), ptr(out[1][1]), ptr(out[1][2])] if L == 2 else [None] * 6
        call("fhvae_lstm_wave_fwd", ptr(P), ptr(Q), ptr(W0), ptr(out[0][0]), ptr(out[0][1]), ptr(out[0][2]), *l1,
             ptr(xchg), T, B, H, L, mode)
        torch.cuda.synchronize()
        for l in range(L):
            for a, b, n in zip(out[l], ref[l], ["h_all", "c_all", "acts"]):
                assert_close(a, b, TC_TOL[mode], f"layer {l} {n} mode {mode} rep {rep}")
    # same launch with the bf16 hi/lo planes of h as extra outputs (TMA operands of the weight-gradient GEMM):
    # fp32 outputs unchanged bit for bit, planes == split of the fp32 h
    out2 = [[f(T, B, H), f(T, B, H), f(T, B, 4 * H)] for _ in range(2)]
    hp = [torch.zeros(2, T * B * H, dtype=torch.bfloat16, device=DEV) for _ in range(2)]
    l1 = [ptr(Wi1), ptr(b1), ptr(W1), ptr(out2[1][0]), ptr(out2[1][1]), ptr(out2[1][2])] if L == 2 else [None] * 6
    # ... and with the weights taken from the pre-packed operand images (fhvae_lstm_wave_pack) instead of being
    # converted in the kernel: same arithmetic, so still bit-identical
    packed = _wave_packed(W0, Wi1 if L == 2 else None, W1 if L == 2 else None, H, L, mode)
    call("fhvae_lstm_wave_fwd_planes", ptr(P), ptr(Q), ptr(W0), ptr(out2[0][0]), ptr(out2[0][1]), ptr(out2[0][2]), *l1,
         ptr(xchg), hp[0].data_ptr(), hp[1].data_ptr() if L == 2 else None, T * B * H, ptr(packed), T, B, H, L, mode)
    torch.cuda.synchronize()
    for l in range(L):
        assert torch.equal(out2[l][0], out[l][0]) and torch.equal(out2[l][2], out[l][2])
        hi = out2[l][0].reshape(-1).to(torch.bfloat16)
        assert torch.equal(hp[l][0], hi)
        assert torch.equal(hp[l][1], (out2[l][0].reshape(-1) - hi.float()).to(torch.bfloat16))


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("T,B,L,use_all,use_last,repeat,H", [(1, 32, 1, True, True, 1, 256), (4, 64, 2, False, True, 2, 256),
                                                             (20, 256, 1, True, False, 2, 256), (20, 256, 2, True, True, 3, 256),
                                                             (20, 256, 2, False, True, 1, 256), (2, 320, 2, True, True, 1, 256),
                                                             (5, 608, 1, True, True, 1, 256),
                                                             (1, 32, 1, True, True, 1, 128), (4, 64, 2, False, True, 2, 128),
                                                             (20, 256, 2, True, True, 3, 128), (20, 256, 1, True, False, 2, 128),
                                                             (3, 608, 2, True, True, 1, 128)])
def test_lstm_wave_bwd_matches_simt(T, B, L, use_all, use_last, repeat, mode, H):
    """BPTT wavefront (top layer first; dh0_t = dg0_{t+1} W_hh0 + dg1_t W_ih1 inside the kernel) against the exact
    fp32 per-layer BPTT kernels + fp32 dgrad GEMM."""
    f = lambda *s: torch.zeros(*s, device=DEV)
    P, Q = rnd(T, B, 4 * H, seed=1, scale=0.7), rnd(B, 4 * H, seed=2, scale=0.3)
    W0 = rnd(4 * H, H, seed=3, scale=1.0 / 16)
    Wi1, W1, b1 = rnd(4 * H, H, seed=4, scale=1.0 / 16), rnd(4 * H, H, seed=5, scale=1.0 / 16), rnd(4 * H, seed=6, scale=0.2)
    st = [[f(T, B, H), f(T, B, H), f(T, B, 4 * H)] for _ in range(2)]       # h, c, acts per layer (exact fp32 forward)
    call("fhvae_lstm_fwd", ptr(P), ptr(Q), ptr(W0), ptr(st[0][0]), ptr(st[0][1]), ptr(st[0][2]), None, T, B, H, 0)
    if L == 2:
        P1 = f(T, B, 4 * H)
        gemm(st[0][0], Wi1, P1, T * B, 4 * H, H, (H, 1), (1, H), 4 * H, bias=b1)
        call("fhvae_lstm_fwd", ptr(P1), None, ptr(W1), ptr(st[1][0]), ptr(st[1][1]), ptr(st[1][2]), None, T, B, H, 0)
    top = L - 1
    Wtop = W1 if L == 2 else W0
    dh_all, dh_last, dh_last0 = rnd(T, B, H, seed=7), rnd(B, H, seed=8), rnd(B, H, seed=9)
    pa = ptr(dh_all) if use_all else None
    pl = ptr(dh_last) if use_last else None
    # reference: per-layer exact kernels
    rdg = [f(T, B, 4 * H), f(T, B, 4 * H)]
    rsum = [f(B, 4 * H), f(B, 4 * H)]
    dc = f(B, H)
    call("fhvae_lstm_bwd", pa, pl, ptr(Wtop), ptr(st[top][1]), ptr(st[top][2]), ptr(rdg[top]), ptr(rsum[top]), None,
         ptr(dc), T, B, H, 0)
    if L == 2:
        dh0 = f(T, B, H)
        gemm(rdg[1], Wi1, dh0, T * B, H, 4 * H, (4 * H, 1), (H, 1), H)      # dh0 = dg1 @ W_ih1
        call("fhvae_lstm_bwd", ptr(dh0), ptr(dh_last0), ptr(W0), ptr(st[0][1]), ptr(st[0][2]), ptr(rdg[0]), ptr(rsum[0]),
             None, ptr(dc), T, B, H, 0)
    n = _lib.fn("fhvae_lstm_wave_bwd_xchg_bytes")(T, B, H, L)
    assert n > 0
    xchg = torch.zeros(n // 4, device=DEV)
    for rep in range(repeat):
        dg = [f(T, B, 4 * H), f(T, B, 4 * H)]
        dgs = [f(B, 4 * H), f(B, 4 * H)]
        bot = ([ptr(Wi1), ptr(W0), ptr(st[0][1]), ptr(st[0][2]), ptr(dg[0]), ptr(dgs[0])] if L == 2 else [None] * 6)
        call("fhvae_lstm_wave_bwd", pa, pl, ptr(dh_last0) if L == 2 else None, ptr(Wtop), ptr(st[top][1]), ptr(st[top][2]),
             ptr(dg[top]), ptr(dgs[top]), *bot, ptr(xchg), T, B, H, L, mode)
        torch.cuda.synchronize()
        for l in range(L):
            assert_close(dg[l], rdg[l], TC_TOL[mode], f"layer {l} dgates mode {mode} rep {rep}")
            assert_close(dgs[l], rsum[l], TC_TOL[mode], f"layer {l} dgsum mode {mode} rep {rep}")
    # planes-only form: fp32 dgates not written at all, planes == split of the fp32 dgates of the plain launch
    dgp = [torch.zeros(2, T * B * 4 * H, dtype=torch.bfloat16, device=DEV) for _ in range(2)]
    dgs2 = [f(B, 4 * H), f(B, 4 * H)]
    bot = ([ptr(Wi1), ptr(W0), ptr(st[0][1]), ptr(st[0][2]), None, ptr(dgs2[0])] if L == 2 else [None] * 6)
    packed = _wave_packed(W0, Wi1 if L == 2 else None, W1 if L == 2 else None, H, L, mode)   # pre-packed weight operands
    call("fhvae_lstm_wave_bwd_planes", pa, pl, ptr(dh_last0) if L == 2 else None, ptr(Wtop), ptr(st[top][1]),
         ptr(st[top][2]), None, ptr(dgs2[top]), *bot, ptr(xchg), dgp[top].data_ptr(),
         dgp[0].data_ptr() if L == 2 else None, T * B * 4 * H, ptr(packed), T, B, H, L, mode)
    torch.cuda.synchronize()
    for l in range(L):
        hi = dg[l].reshape(-1).to(torch.bfloat16)
        assert torch.equal(dgp[l][0], hi), f"layer {l} dgates hi plane"
        assert torch.equal(dgp[l][1], (dg[l].reshape(-1) - hi.float()).to(torch.bfloat16)), f"layer {l} dgates lo plane"
        assert torch.equal(dgs2[l], dgs[l])


def test_lstm_null_inputs():
    T, B, H = 3, 8, 16
    f = lambda *s: torch.zeros(*s, device=DEV)
    Q, W = rnd(B, 4 * H, seed=1), rnd(4 * H, H, seed=2, scale=0.2)
    h1, c1, a1 = f(T, B, H), f(T, B, H), f(T, B, 4 * H)
    call("fhvae_lstm_fwd", None, ptr(Q), ptr(W), ptr(h1), ptr(c1), ptr(a1), None, T, B, H, 0)
    P = Q.unsqueeze(0).expand(T, B, 4 * H).contiguous()
    h2, c2, a2 = f(T, B, H), f(T, B, H), f(T, B, 4 * H)
    call("fhvae_lstm_fwd", ptr(P), None, ptr(W), ptr(h2), ptr(c2), ptr(a2), None, T, B, H, 0)
    assert torch.equal(h1, h2) and torch.equal(c1, c2)
    assert _lib.fn("fhvae_lstm_fwd")(None, None, ptr(W), ptr(h1), ptr(c1), ptr(a1), None, T, B, H, 0, None) == -1


# ------------------------------------------------------------------------------- reparam / ELBO
@pytest.mark.parametrize("B,T,F,Z", [(7, 4, 6, 16), (256, 20, 80, 32), (64, 20, 80, 16)])
@pytest.mark.parametrize("layout", ["time_major", "simple"])
def test_elbo_fwd_bwd(B, T, F, Z, layout):
    g = torch.Generator().manual_seed(B + T)
    x = torch.randn(B, T, F, generator=g)
    xm = (torch.randn(B, T, F, generator=g) * 0.5).requires_grad_(True)
    xl = (torch.randn(B, T, F, generator=g) * 0.3).requires_grad_(True)
    z1h = (torch.randn(B, 2 * Z, generator=g) * 0.5).requires_grad_(True)
    z2h = (torch.randn(B, 2 * Z, generator=g) * 0.5).requires_grad_(True)
    mu2 = torch.randn(B, Z, generator=g).requires_grad_(True)
    nsegs = torch.randint(1, 200, (B,), generator=g)
    lb, lpx, nk1, nk2, lpm = O.elbo_terms(x, xm, xl, z1h[:, :Z], z1h[:, Z:], z2h[:, :Z], z2h[:, Z:], mu2, nsegs)
    w = torch.randn(5, B, generator=g)
    (w[0] * lb + w[1] * lpx + w[2] * nk1 + w[3] * nk2 + w[4] * lpm).sum().backward()
    if layout == "time_major":      # (T,B,2F): [mu | logvar]
        xhead = torch.cat([xm, xl], -1).permute(1, 0, 2).contiguous().detach().to(DEV)
        xs_b, xs_t, lv_off = 2 * F, B * 2 * F, F
        unpack = lambda d: (d.permute(1, 0, 2)[..., :F], d.permute(1, 0, 2)[..., F:])
    else:                           # (B, 2TF): [mu(TF) | logvar(TF)]
        xhead = torch.cat([xm.reshape(B, -1), xl.reshape(B, -1)], -1).contiguous().detach().to(DEV)
        xs_b, xs_t, lv_off = 2 * T * F, F, T * F
        unpack = lambda d: (d[:, :T * F].view(B, T, F), d[:, T * F:].view(B, T, F))
    d = lambda t: t.detach().to(DEV).contiguous()
    out5 = torch.zeros(5, B, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    xd, z1d, z2d, mu2d, nsd = d(x), d(z1h), d(z2h), d(mu2), d(nsegs)
    call("fhvae_elbo_fwd", ptr(xd), ptr(xhead), xs_b, xs_t, lv_off, ptr(z1d), ptr(z2d), ptr(mu2d),
         ptr(nsd), ptr(out5), ptr(flag), B, T, F, Z, Z)
    for i, r in enumerate([lb, lpx, nk1, nk2, lpm]):
        assert_close(out5[i], r, 1e-5, f"elbo out {i}")
    assert int(flag) == 0
    # coef as the host folds it: c_px = w_lb + w_px ... c_pm = w_lb/nsegs + w_pm
    coef = torch.stack([w[0] + w[1], w[0] + w[2], w[0] + w[3], w[0] / nsegs + w[4]]).to(DEV).contiguous()
    dxh = torch.zeros_like(xhead)
    dz1, dz2, dmu2 = torch.zeros(B, 2 * Z, device=DEV), torch.zeros(B, 2 * Z, device=DEV), torch.zeros(B, Z, device=DEV)
    call("fhvae_elbo_bwd", ptr(xd), ptr(xhead), xs_b, xs_t, lv_off, ptr(z1d), ptr(z2d), ptr(mu2d),
         ptr(coef), ptr(dxh), ptr(dz1), ptr(dz2), ptr(dmu2), B, T, F, Z, Z)
    dm, dl = unpack(dxh)
    assert_close(dm, xm.grad, 1e-5, "d x_mu"); assert_close(dl, xl.grad, 1e-5, "d x_logvar")
    assert_close(dz1, z1h.grad, 1e-5, "d z1head"); assert_close(dz2, z2h.grad, 1e-5, "d z2head")
    assert_close(dmu2, mu2.grad, 1e-5, "d mu2")


def test_elbo_fwd_bwd_fused_is_bit_identical():
    """fhvae_elbo_fwd_bwd == fhvae_elbo_fwd + fhvae_step_coef + fhvae_elbo_bwd, bit for bit (both decoder-head layouts)."""
    B, T, F, Z = 37, 6, 8, 12
    for layout in ("tbf", "btf"):
        x = rnd(B, T, F, seed=1)
        xh = rnd(T, B, 2 * F, seed=2) if layout == "tbf" else rnd(B, 2 * T * F, seed=2)
        xs_b, xs_t, lv = (2 * F, B * 2 * F, F) if layout == "tbf" else (2 * T * F, F, T * F)
        h1, h2, m2 = rnd(B, 2 * Z, seed=3), rnd(B, 2 * Z, seed=4), rnd(B, Z, seed=5)
        ns = torch.randint(1, 300, (B,), generator=torch.Generator().manual_seed(6)).to(DEV)
        gout = rnd(6, B, seed=7)
        for detach, prior in ((0, 1), (1, 0)):
            out_a, out_b = torch.zeros(5, B, device=DEV), torch.zeros(5, B, device=DEV)
            fl = torch.zeros(1, dtype=torch.int32, device=DEV)
            coef = torch.zeros(4, B, device=DEV)
            d_a = [torch.zeros_like(xh), torch.zeros(B, 2 * Z, device=DEV), torch.zeros(B, 2 * Z, device=DEV), torch.zeros(B, Z, device=DEV)]
            d_b = [torch.zeros_like(t) for t in d_a]
            call("fhvae_elbo_fwd", ptr(x), ptr(xh), xs_b, xs_t, lv, ptr(h1), ptr(h2), ptr(m2), ptr(ns), ptr(out_a), ptr(fl),
                 B, T, F, Z, Z)
            call("fhvae_step_coef", ptr(gout), ptr(ns), ptr(coef), detach, prior, B)
            call("fhvae_elbo_bwd", ptr(x), ptr(xh), xs_b, xs_t, lv, ptr(h1), ptr(h2), ptr(m2), ptr(coef), ptr(d_a[0]),
                 ptr(d_a[1]), ptr(d_a[2]), ptr(d_a[3]), B, T, F, Z, Z)
            call("fhvae_elbo_fwd_bwd", ptr(x), ptr(xh), xs_b, xs_t, lv, ptr(h1), ptr(h2), ptr(m2), ptr(ns), ptr(gout), detach,
                 prior, ptr(out_b), ptr(fl), ptr(d_b[0]), ptr(d_b[1]), ptr(d_b[2]), ptr(d_b[3]), B, T, F, Z, Z)
            assert torch.equal(out_a, out_b)
            for a, b in zip(d_a, d_b):
                assert torch.equal(a, b)


def test_elbo_nan_flag():
    B, T, F, Z = 4, 2, 4, 8
    z = lambda *s: torch.zeros(*s, device=DEV)
    x = z(B, T, F); x[2, 0, 0] = float("nan")
    out5, flag = z(5, B), torch.zeros(1, dtype=torch.int32, device=DEV)
    xh, h1, h2, m2, ns = z(T, B, 2 * F), z(B, 2 * Z), z(B, 2 * Z), z(B, Z), torch.ones(B, dtype=torch.int64, device=DEV)
    call("fhvae_elbo_fwd", ptr(x), ptr(xh), 2 * F, B * 2 * F, F, ptr(h1), ptr(h2),
         ptr(m2), ptr(ns), ptr(out5), ptr(flag), B, T, F, Z, Z)
    assert int(flag) == 1 and torch.isnan(out5[0, 2]) and not torch.isnan(out5[0, 0])


def test_reparam_fwd_bwd():
    B, Z = 33, 16
    head, eps, ds = rnd(B, 2 * Z, seed=1), rnd(B, Z, seed=2), rnd(B, Z + 5, seed=3)
    s = torch.zeros(B, Z + 3, device=DEV)
    call("fhvae_reparam_fwd", ptr(head), 2 * Z, ptr(eps), ptr(s, 3), Z + 3, B, Z)
    ref = head[:, :Z] + eps * torch.exp(0.5 * head[:, Z:])
    assert_close(s[:, 3:], ref, 1e-6, "reparam")
    dh = torch.ones(B, 2 * Z, device=DEV)
    call("fhvae_reparam_bwd", ptr(head), 2 * Z, ptr(eps), ptr(ds, 5), Z + 5, ptr(dh), 2 * Z, 1, B, Z)
    g = ds[:, 5:]
    assert_close(dh[:, :Z], 1 + g, 1e-6, "dmu"); assert_close(dh[:, Z:], 1 + g * 0.5 * eps * torch.exp(0.5 * head[:, Z:]), 1e-6, "dlv")


# ------------------------------------------------------------------------------- fused latent-head stages
@pytest.mark.parametrize("B,H,L,Z,Kq,NQ,with_q", [(256, 256, 2, 32, 32, 1024, True), (9, 32, 2, 8, 24, 128, True),
                                                   (33, 64, 1, 16, 16, 256, True), (5, 40, 2, 12, 0, 0, False)])
def test_head_fwd_matches_fp64(B, H, L, Z, Kq, NQ, with_q):
    """head = [h_l0 | h_l1] W^T + b ; z = mu + eps exp(lv/2) ; Q = zcat[:, qoff:qoff+Kq] Wq^T + bq
    (GaussianLayer, simple_fhvae.py:205-216, + the hoisted projection of the next stack)."""
    hs = [rnd(B, H, seed=10 + l) for l in range(L)]
    W, b, eps = rnd(2 * Z, L * H, seed=1, scale=0.1), rnd(2 * Z, seed=2), rnd(B, Z, seed=3)
    ldz, zoff = Z + 24, 24                               # sample lands in zcat[:, 24:24+Z]; Q reads zcat[:, qoff:]
    zcat = rnd(B, ldz, seed=4)
    zc0 = zcat.clone()
    qoff = ldz - Kq
    ldw = Kq + 7
    Wq, bq = rnd(max(NQ, 1), ldw, seed=5), rnd(max(NQ, 1), seed=6)
    head = torch.zeros(B, 2 * Z, device=DEV)
    Q = torch.zeros(B, max(NQ, 1), device=DEV)
    call("fhvae_head_fwd", ptr(hs[0]), ptr(hs[1]) if L > 1 else None, H, L, H, ptr(W), ptr(b), ptr(head), Z,
         ptr(eps), ptr(zcat), ldz, zoff, ptr(Wq, 3) if with_q else None, ldw, ptr(bq) if with_q else None, qoff, Kq,
         ptr(Q) if with_q else None, NQ, B)
    hc = torch.cat(hs, 1).double()
    href = hc @ W.double().t() + b.double()
    assert_close(head, href, 3e-6, "head")
    zref = href[:, :Z] + eps.double() * torch.exp(0.5 * href[:, Z:])
    assert_close(zcat[:, zoff:], zref, 3e-6, "sample")
    assert torch.equal(zcat[:, :zoff], zc0[:, :zoff])    # nothing outside the sample slice is touched
    if with_q:
        zc = zc0.double().clone(); zc[:, zoff:] = zref
        qref = zc[:, qoff:qoff + Kq] @ Wq[:, 3:3 + Kq].double().t() + bq.double()
        assert_close(Q, qref, 3e-6, "Q")


@pytest.mark.parametrize("B,H,L,Z,NG,Kq,beta", [(256, 256, 2, 32, 1024, 64, 0), (256, 256, 2, 32, 1024, 32, 1),
                                                 (9, 32, 2, 8, 128, 8, 1), (33, 64, 1, 16, 256, 40, 0)])
def test_head_bwd_matches_fp64(B, H, L, Z, NG, Kq, beta):
    dgsum, ldw = rnd(B, NG, seed=1), Kq + 5
    Wq = rnd(NG, ldw, seed=2, scale=0.1)
    ldz = max(Kq, Z) + 16
    dzoff = ldz - Kq
    roff = ldz - Z                                       # reparam slice inside the freshly written columns
    dzcat = rnd(B, ldz, seed=3)
    dz0 = dzcat.clone()
    head, eps, dhead = rnd(B, 2 * Z, seed=4), rnd(B, Z, seed=5), rnd(B, 2 * Z, seed=6)
    dh0 = dhead.clone()
    W = rnd(2 * Z, L * H, seed=7, scale=0.1)
    dh = [torch.zeros(B, H, device=DEV) for _ in range(L)]
    call("fhvae_head_bwd", ptr(dgsum), NG, ptr(Wq, 2), ldw, Kq, ptr(dzcat), ldz, dzoff, beta, ptr(head), ptr(eps), Z,
         roff, ptr(dhead), 1, ptr(W), L, H, ptr(dh[0]), ptr(dh[1]) if L > 1 else None, B)
    dzr = dz0.double().clone()
    upd = dgsum.double() @ Wq[:, 2:2 + Kq].double()
    dzr[:, dzoff:] = upd + (dzr[:, dzoff:] if beta else 0)
    assert_close(dzcat, dzr, 3e-6, "dzcat")
    gz = dzr[:, roff:roff + Z]
    dhr = dh0.double().clone()
    dhr[:, :Z] += gz
    dhr[:, Z:] += gz * 0.5 * eps.double() * torch.exp(0.5 * head[:, Z:].double())
    assert_close(dhead, dhr, 3e-6, "dhead")
    full = dhr @ W.double()
    for l in range(L):
        assert_close(dh[l], full[:, l * H:(l + 1) * H], 3e-6, f"dh_last[{l}]")
    # head-only form (detach_px path): no dz, no reparam
    dh2 = [torch.zeros(B, H, device=DEV) for _ in range(L)]
    call("fhvae_head_bwd", None, 0, None, 0, 0, None, 0, 0, 0, None, None, Z, 0, ptr(dhead), 1, ptr(W), L, H,
         ptr(dh2[0]), ptr(dh2[1]) if L > 1 else None, B)
    for l in range(L):
        assert torch.equal(dh2[l], dh[l])


def test_step_coef_and_loss_mean():
    B = 77
    gout = rnd(6, B, seed=1)
    ns = torch.randint(1, 200, (B,), generator=torch.Generator().manual_seed(2)).to(DEV)
    coef = torch.zeros(4, B, device=DEV)
    for detach, prior in ((0, 1), (1, 0)):
        call("fhvae_step_coef", ptr(gout), ptr(ns), ptr(coef), detach, prior, B)
        ref = torch.stack([gout[1] + gout[0], gout[2] + gout[0], gout[3] + gout[0], gout[0] / ns.float() + gout[4]])
        if detach:
            ref[0] = 0
        if not prior:
            ref[3] = 0
        assert torch.equal(coef, ref)
    loss = torch.zeros((), device=DEV)
    call("fhvae_loss_mean", ptr(gout), ptr(gout, 5 * B), 10.0, B, ptr(loss))
    assert_close(loss.reshape(1), -(gout[0].double() + 10.0 * gout[5].double()).mean().reshape(1), 1e-6, "loss")


def test_load_inputs_exact():
    """x / mu_idx / num_segs into the static buffers in one launch, x also time-major: bit-exact copies."""
    B, T, F = 37, 5, 8
    x = rnd(B, T, F, seed=1)
    g = torch.Generator().manual_seed(2)
    idx, ns = torch.randint(0, 1000, (B,), generator=g).to(DEV), torch.randint(1, 99, (B,), generator=g).to(DEV)
    xd, xtm = torch.zeros(B, T, F, device=DEV), torch.zeros(T, B, F, device=DEV)
    idd, nsd = torch.zeros(B, dtype=torch.int64, device=DEV), torch.zeros(B, dtype=torch.int64, device=DEV)
    call("fhvae_load_inputs", ptr(x), ptr(xd), ptr(xtm), B, T, F, ptr(idx), ptr(idd), ptr(ns), ptr(nsd))
    assert torch.equal(xd, x) and torch.equal(xtm, x.permute(1, 0, 2)) and torch.equal(idd, idx) and torch.equal(nsd, ns)
    xd2 = torch.zeros(B, T, F, device=DEV)
    call("fhvae_load_inputs", ptr(x), ptr(xd2), None, B, T, F, None, None, None, None)
    assert torch.equal(xd2, x)


# ------------------------------------------------------------------------------- discriminative term
@pytest.mark.parametrize("B,N,Z", [(5, 12, 16), (64, 1000, 16), (256, 5000, 32), (33, 129, 8)])
def test_disc_fwd_bwd(B, N, Z):
    g = torch.Generator().manual_seed(N)
    table = torch.randn(N, Z, generator=g).double().requires_grad_(True)
    idx = torch.randint(0, N, (B,), generator=g)
    idx[0], idx[-1] = 0, N - 1
    if B > 3:
        idx[2] = idx[1]                                  # duplicate utterance in the batch
    z2h = torch.randn(B, 2 * Z, generator=g).double()
    # posterior means pulled towards their rows but with a non-degenerate softmax (log q ~ -1..-10)
    z2h[:, :Z] = 0.35 * table.detach()[idx] + 0.25 * z2h[:, :Z]
    z2h.requires_grad_(True)
    lq = O.log_qy_per_segment(z2h[:, :Z], table, idx)
    w = torch.randn(B, generator=g).double()
    (w * lq).sum().backward()
    d = lambda t: t.detach().float().to(DEV).contiguous()
    zd, td, idd = d(z2h), d(table), idx.to(DEV)
    ns = _lib.fn("fhvae_disc_nsplit")(B, N)
    f = lambda *s: torch.zeros(*s, device=DEV)
    part, tgt, lqd, lse, mu2 = f(ns, B, 2), f(B), f(B), f(B), f(B, Z)
    call("fhvae_mu2_gather", ptr(td), ptr(idd), ptr(mu2), B, Z, N, None)
    assert torch.equal(mu2, td[idd])
    call("fhvae_disc_fwd_partial", ptr(zd), 2 * Z, ptr(td), N, Z, ptr(part), ns, B)
    call("fhvae_disc_target", ptr(zd), 2 * Z, ptr(mu2), ptr(tgt), B, Z)
    call("fhvae_disc_combine", ptr(part), ns, ptr(tgt), ptr(lqd), ptr(lse), B)
    assert_close(lqd, lq, 2e-5, "log_qy")
    gq = d(w)
    dtab, sumpm, dz, dmu2 = f(N, Z), f(ns, B, Z), f(B, 2 * Z), f(B, Z)
    call("fhvae_disc_bwd_segs", ptr(zd), 2 * Z, ptr(td), N, Z, ptr(lse), ptr(sumpm), ns, B)
    call("fhvae_disc_bwd_rows", ptr(zd), 2 * Z, ptr(td), N, Z, ptr(lse), ptr(gq), ptr(dtab), B)
    call("fhvae_disc_bwd_finish", ptr(zd), 2 * Z, ptr(mu2), ptr(sumpm), ns, ptr(gq), ptr(dz), 2 * Z, ptr(dmu2), B, Z)
    touched = torch.zeros(B, dtype=torch.int32, device=DEV)
    call("fhvae_mu2_scatter_reduce", ptr(dmu2), ptr(idd), ptr(dtab), ptr(touched), B, Z, N)
    assert_close(dz[:, :Z], z2h.grad[:, :Z], 5e-5, "d z2_mu")
    assert float(dz[:, Z:].abs().max()) == 0.0
    assert_close(dtab, table.grad, 5e-5, "d table")
    # rows touched by the sparse part: exactly the distinct utterances, flagged at first occurrence
    first = torch.zeros(B, dtype=torch.int32)
    seen = set()
    for b, r in enumerate(idx.tolist()):
        if r not in seen:
            first[b] = 1; seen.add(r)
    assert torch.equal(touched.cpu(), first)


def test_scatter_reduce_exact_and_deterministic():
    B, Z, N = 512, 32, 40                                  # heavy duplication
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, N, (B,), generator=g)
    src = torch.randint(-8, 9, (B, Z), generator=g).float()   # small integers: float sums are exact
    ref = torch.zeros(N, Z).index_add_(0, idx, src)
    outs = []
    for _ in range(2):
        dst = torch.zeros(N, Z, device=DEV)
        sd, idd = src.to(DEV), idx.to(DEV)
        call("fhvae_mu2_scatter_reduce", ptr(sd), ptr(idd), ptr(dst), None, B, Z, N)
        outs.append(dst.cpu())
    assert torch.equal(outs[0], ref) and torch.equal(outs[0], outs[1])


def test_mu2_estimate_matches_reference_dict_loop():
    B, Z, K = 300, 32, 17
    g = torch.Generator().manual_seed(5)
    z2h = torch.randn(B, 2 * Z, generator=g)
    idx = torch.randint(0, K - 2, (B,), generator=g)       # rows K-2, K-1 never seen: stay untouched
    d = O.estimate_mu2_dict([z2h[:100, :Z], z2h[100:, :Z]], [idx[:100], idx[100:]])
    zsum, cnt = torch.zeros(K, Z, device=DEV), torch.zeros(K, device=DEV)
    table = torch.full((K, Z), 7.0, device=DEV)
    zd, idd = z2h.to(DEV), idx.to(DEV)
    call("fhvae_mu2_accumulate", ptr(zd), 2 * Z, ptr(idd), ptr(zsum), ptr(cnt), 100, Z, K, None)
    call("fhvae_mu2_accumulate", ptr(zd, 100 * 2 * Z), 2 * Z, ptr(idd, 100), ptr(zsum), ptr(cnt), B - 100, Z, K, None)
    call("fhvae_mu2_estimate_finish", ptr(zsum), ptr(cnt), ptr(table), 0.25, K, Z)
    for y, v in d.items():
        assert_close(table[y], v, 1e-5, f"mu2[{y}]")
    assert torch.equal(cnt.cpu().long(), torch.bincount(idx, minlength=K))
    assert float((table[K - 2:] - 7.0).abs().max()) == 0.0


def test_rows_copy_exact():
    src = rnd(50, 16, seed=1)
    dst = torch.zeros(20, 16, device=DEV)
    s_rows = torch.tensor([49, 0, 7, 7], device=DEV)
    d_rows = torch.tensor([0, 19, -1, 3], device=DEV)
    call("fhvae_rows_copy", ptr(src), ptr(s_rows), ptr(dst), ptr(d_rows), 4, 16)
    assert torch.equal(dst[0], src[49]) and torch.equal(dst[19], src[0]) and torch.equal(dst[3], src[7])
    assert float(dst[1:3].abs().max()) == 0.0


# ------------------------------------------------------------------------------- Adam / helpers
def test_adam_flat_matches_torch_adam():
    n = 100003
    p0, gs = rnd(n, seed=1), [rnd(n, seed=10 + i) for i in range(4)]
    ref = torch.nn.Parameter(p0.clone().cpu())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.95, 0.999))
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    done = torch.zeros(1, dtype=torch.int32, device=DEV)
    for g in gs:
        ref.grad = g.cpu().clone(); opt.step()
        call("fhvae_adam_flat", ptr(p), ptr(g), ptr(m), ptr(v), n, 1e-3, 0.95, 0.999, 1e-8, 1.0, ptr(step), ptr(done))
    assert int(step) == 4 and int(done) == 0
    assert_close(p, ref.data, 1e-6, "adam params")
    assert_close(m, opt.state[ref]["exp_avg"], 1e-6, "adam m")


def test_transpose_colsum_add2_relu():
    B, T, F = 9, 5, 12
    x = rnd(B, T, F, seed=1)
    y = torch.zeros(T, B, F, device=DEV)
    call("fhvae_transpose_bt", ptr(x), ptr(y), B, T, F)
    assert torch.equal(y, x.permute(1, 0, 2).contiguous())
    a = rnd(777, 70, seed=2)
    o1, o2 = torch.zeros(70, device=DEV), torch.zeros(70, device=DEV)
    call("fhvae_colsum_batch", (ColsumProblem * 1)(ColsumProblem(ptr(a), ptr(o1), ptr(o2), 70, 777, 70)), 1)
    assert_close(o1, a.double().sum(0), 1e-5, "colsum"); assert torch.equal(o1, o2)
    s = torch.zeros(777 * 70, device=DEV)
    call("fhvae_add2", ptr(s), ptr(a), ptr(a), 777 * 70)
    assert torch.equal(s.view(777, 70), a + a)
    d = torch.ones_like(a)
    call("fhvae_relu_bwd", ptr(d), ptr(a), a.numel())
    assert torch.equal(d, (a > 0).float())


def test_randn_kernel_statistics_and_stream_advance():
    """fhvae_randn (the step's own eps draw): N(0,1) moments, no repeats across launches / graph replays, reproducible
    from (seed, offset)."""
    n = 1 << 20
    out = torch.zeros(n + 3, device=DEV)
    st = torch.zeros(2, dtype=torch.int64, device=DEV)
    call("fhvae_randn", ptr(out), n + 3, 1234, ptr(st), ptr(st, 1))
    a = out.clone()
    assert int(st[0]) == (n + 3 + 3) // 4 and int(st[1]) == 0
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.var()) - 1.0) < 5e-3
    assert abs(float((a ** 3).mean())) < 2e-2 and abs(float((a ** 4).mean()) - 3.0) < 5e-2
    assert float(a.abs().max()) < 7.0 and bool(torch.isfinite(a).all())
    call("fhvae_randn", ptr(out), n + 3, 1234, ptr(st), ptr(st, 1))
    b = out.clone()
    assert float((a == b).float().mean()) < 1e-3                                   # the stream advanced
    assert abs(float((a * b).mean())) < 5e-3                                       # ... and is uncorrelated
    st2 = torch.zeros(2, dtype=torch.int64, device=DEV)
    call("fhvae_randn", ptr(out), n + 3, 1234, ptr(st2), ptr(st2, 1))
    assert torch.equal(out, a)                                                     # same (seed, offset) -> same draws
    call("fhvae_randn", ptr(out), n + 3, 99, ptr(torch.zeros(2, dtype=torch.int64, device=DEV)), ptr(st2, 1))
    assert float((out == a).float().mean()) < 1e-3                                 # another seed -> another stream
    # graph replays draw fresh numbers (the offset lives on the device)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=s):
        _lib.check(_lib.fn("fhvae_randn")(ptr(out), n, 7, ptr(st), ptr(st, 1), torch.cuda.current_stream().cuda_stream))
    g.replay(); c = out.clone(); g.replay()
    assert float((c == out).float().mean()) < 1e-3


def _kplanes(t):
    """bf16 hi/lo planes [2][rows*ld] of a (rows, K) fp32 tensor (what fhvae_split_planes_batch writes)."""
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return torch.stack([hi.reshape(-1), lo.reshape(-1)]).contiguous()


@pytest.mark.parametrize("mode,tol", [(1, 2e-5), (2, 1.5e-2)])
@pytest.mark.parametrize("M,N,K", [(5120, 1024, 80), (5120, 160, 256), (200, 72, 16), (128, 128, 32), (1, 8, 48), (333, 200, 112)])
def test_proj_planes_gemm_matches_fp64(M, N, K, mode, tol):
    """fhvae_proj_planes_batch (TMA-fed K-major tcgen05 GEMM with the coalesced epilogue) against fp64: the layer-0 input
    projection (5120 x 1024 x 80), the decoder head (5120 x 160 x 256) and ragged M / N / K tails."""
    from pytorch_scalablefhvae_b200._lib import ProjProblem
    A, W, bias = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.2), rnd(N, seed=3)
    pa, pw = _kplanes(A), _kplanes(W)
    ldc = N + 8
    C = torch.full((M, ldc), 7.0, device=DEV)
    prob = ProjProblem(pa.data_ptr(), pw.data_ptr(), ptr(C), ptr(bias), M, N, K, 0, K, M * K, K, N * K, ldc)
    call("fhvae_proj_planes_batch", (ProjProblem * 1)(prob), 1, mode)
    ref = A.double() @ W.double().t() + bias.double()
    assert_close(C[:, :N], ref, tol, f"proj {M}x{N}x{K} mode {mode}")
    assert float((C[:, N:] - 7.0).abs().max()) == 0.0                      # nothing written beyond N
    # two problems in one launch, one without bias
    C1, C2 = torch.zeros(M, N, device=DEV), torch.zeros(M, N, device=DEV)
    arr = (ProjProblem * 2)(ProjProblem(pa.data_ptr(), pw.data_ptr(), ptr(C1), ptr(bias), M, N, K, 0, K, M * K, K, N * K, N),
                            ProjProblem(pa.data_ptr(), pw.data_ptr(), ptr(C2), None, M, N, K, 0, K, M * K, K, N * K, N))
    call("fhvae_proj_planes_batch", arr, 2, mode)
    assert torch.equal(C1, C[:, :N].contiguous()) and torch.equal(C2 + bias, C1)
