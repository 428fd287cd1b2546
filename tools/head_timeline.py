"""Phase stamps (SM clocks) of head_fwd_kernel, CTA 0 / thread 0 (-DFHVAE_TIMELINE build).  Development aid."""
import ctypes, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_scalablefhvae_b200 import _lib
from pytorch_scalablefhvae_b200.plan import ptr

so = "/tmp/libfhvae_tl.so"
srcs = [os.path.join(_lib.CSRC, s) for s in _lib.SOURCES]
subprocess.check_call(["nvcc"] + _lib.NVCC_FLAGS + ["-DFHVAE_TIMELINE", "-o", so] + srcs)
lib = ctypes.CDLL(so)
dev = "cuda"
B, H, L, Z, Kq, NQ = 256, 256, 2, 32, 32, 1024
hs = [torch.randn(B, H, device=dev) for _ in range(L)]
W, b, eps = torch.randn(2 * Z, L * H, device=dev), torch.randn(2 * Z, device=dev), torch.randn(B, Z, device=dev)
zcat = torch.randn(B, 64, device=dev)
Wq = torch.randn(NQ, 112, device=dev)
head, Q = torch.zeros(B, 2 * Z, device=dev), torch.zeros(B, NQ, device=dev)
vp = lambda t: ctypes.c_void_p(ptr(t))
for it in range(3):
    r = lib.fhvae_head_fwd(vp(hs[0]), vp(hs[1]), ctypes.c_int64(H), L, H, vp(W), vp(b), vp(head), Z, vp(eps), vp(zcat), ctypes.c_int64(64), 32,
                           vp(Wq), ctypes.c_int64(112), None, 32, Kq, vp(Q), NQ, B, None)
    assert r == 0, r
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 32)()
    lib.fhvae_debug_head_timeline(buf)
    t = list(buf)[:8]
    print("fwd: stage_h", t[1] - t[0], "stage_W", t[2] - t[1], "head", t[3] - t[2], "store+sample", t[4] - t[3], "stage_Wq", t[5] - t[4], "Q", t[6] - t[5], "total", t[6] - t[0])
