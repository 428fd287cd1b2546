"""Build a -DFHVAE_TIMELINE copy of the library and print the per-phase SM-clock deltas of one LSTM
forward call (CTA 0, thread 0).  Development aid."""
import ctypes, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_scalablefhvae_b200 import _lib
from pytorch_scalablefhvae_b200.plan import ptr

so = "/tmp/libfhvae_tl.so"
srcs = [os.path.join(_lib.CSRC, s) for s in _lib.SOURCES]
subprocess.check_call(["nvcc"] + _lib.NVCC_FLAGS + ["-DFHVAE_TIMELINE", "-o", so] + srcs)
lib = ctypes.CDLL(so)
T, B, H = 20, 256, 256
z = lambda *s: torch.randn(*s, device="cuda") * 0.3
P, Q, W = z(T, B, 4 * H), z(B, 4 * H), z(4 * H, H) / 16
h, c, a = z(T, B, H), z(T, B, H), z(T, B, 4 * H)
xch = z(16, B, H)
names = ["top", "cl_wait", "mma_issued", "mma_done", "tmem_ld", "gates", "sync1", "cell", "sync2", "dsmem_st", "arrive", "hbm_st"]
for mode in (1, 2):
    for _ in range(3):
        lib.fhvae_lstm_fwd(ctypes.c_void_p(ptr(P)), ctypes.c_void_p(ptr(Q)), ctypes.c_void_p(ptr(W)), ctypes.c_void_p(ptr(h)),
                           ctypes.c_void_p(ptr(c)), ctypes.c_void_p(ptr(a)), ctypes.c_void_p(ptr(xch)), T, B, H, mode, None)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (2 * 32 * 16))()
    lib.fhvae_debug_timeline(buf)
    tl = [[buf[(0 * 32 + t) * 16 + k] for k in range(16)] for t in range(T)]
    print(f"mode {mode}: cycles per phase (steps 5..9), step period:")
    for t in range(5, 10):
        d = [tl[t][k] - tl[t][k - 1] for k in range(1, 12)]
        print(t, dict(zip(names[1:], d)), "fence", tl[t][12] - tl[t][1], "period", tl[t][0] - tl[t - 1][0])

# ---- backward kernel
dg, dgs, dhr, dc, dh = z(T, B, 4 * H), z(B, 4 * H), z(16, B, H), z(B, H), z(T, B, H)
bn = ["top", "cl_wait", "sum_recv", "pointwise", "fence_sync", "mma_issued", "mma_done", "tmem+dsmem_st", "arrive"]
for mode in (1, 2):
    for _ in range(3):
        lib.fhvae_lstm_bwd(ctypes.c_void_p(ptr(dh)), None, ctypes.c_void_p(ptr(W)), ctypes.c_void_p(ptr(c)), ctypes.c_void_p(ptr(a)),
                           ctypes.c_void_p(ptr(dg)), ctypes.c_void_p(ptr(dgs)), ctypes.c_void_p(ptr(dhr)), ctypes.c_void_p(ptr(dc)), T, B, H, mode, None)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (2 * 32 * 16))()
    lib.fhvae_debug_timeline(buf)
    tl = [[buf[(1 * 32 + t) * 16 + k] for k in range(16)] for t in range(T)]
    print(f"bwd mode {mode}:")
    for t in range(5, 8):
        d = [tl[t][k] - tl[t][k - 1] for k in range(1, 9)]
        print(t, dict(zip(bn[1:], d)), "period", tl[t][0] - tl[t - 1][0])
