"""Per-phase SM-clock stamps of the wavefront LSTM forward (-DFHVAE_TIMELINE build): group 0 / CTA 0 / thread 0
of each layer.  Development aid."""
import ctypes, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_scalablefhvae_b200 import _lib
from pytorch_scalablefhvae_b200.plan import ptr

so = "/tmp/libfhvae_tl.so"
srcs = [os.path.join(_lib.CSRC, s) for s in _lib.SOURCES]
subprocess.check_call(["nvcc"] + _lib.NVCC_FLAGS + ["-DFHVAE_TIMELINE", "-o", so] + srcs)
lib = ctypes.CDLL(so)
lib.fhvae_lstm_wave_xchg_bytes.restype = ctypes.c_longlong
T, B, H = 20, 256, 256
z = lambda *s: torch.randn(*s, device="cuda") * 0.3
P, Q = z(T, B, 4 * H), z(B, 4 * H)
W0, Wi1, W1, b1 = z(4 * H, H) / 16, z(4 * H, H) / 16, z(4 * H, H) / 16, z(4 * H)
o = [[z(T, B, H), z(T, B, H), z(T, B, 4 * H)] for _ in range(2)]
vp = lambda t: ctypes.c_void_p(ptr(t))
names = ["top", "pull+hbm_st+arrive(+P prefetch)", "rec_done", "tmem_ld", "gates+sync", "cell+sync", "publish", "p1_readout"]
for L in (1, 2):
    xchg = torch.zeros(lib.fhvae_lstm_wave_xchg_bytes(T, B, H, L) // 4, device="cuda")
    for mode in (1, 2):
        for _ in range(3):
            l1 = [vp(Wi1), vp(b1), vp(W1), vp(o[1][0]), vp(o[1][1]), vp(o[1][2])] if L == 2 else [None] * 6
            r = lib.fhvae_lstm_wave_fwd(vp(P), vp(Q), vp(W0), vp(o[0][0]), vp(o[0][1]), vp(o[0][2]), *l1, vp(xchg), T, B, H, L, mode, None)
            assert r == 0
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * (2 * 32 * 16))()
        lib.fhvae_debug_wave_timeline(buf)
        for layer in range(L):
            tl = [[buf[(layer * 32 + t) * 16 + k] for k in range(16)] for t in range(T)]
            print(f"L={L} mode {mode} layer {layer}: cycles per phase (steps 8..11)")
            for t in range(8, 12):
                r = tl[t]
                print("  ", t, "loads_issued", r[6] - r[0], "waited+stored", r[7] - r[6], "pull_total", r[1] - r[0], "rec_done", r[2] - r[1], "tmem_ld", r[3] - r[2],
                      "gates", r[4] - r[3], "cell", r[5] - r[4], "rest(publish,p1)", tl[t + 1][0] - r[5], "period", r[0] - tl[t - 1][0])
        if L == 2:
            print("   layer-1 lag behind layer 0 at step 10 top (cycles; clocks of different SMs, indicative):",
                  buf[(32 + 10) * 16] - buf[10 * 16])
