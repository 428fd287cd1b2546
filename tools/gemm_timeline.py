"""Per-phase SM-clock stamps of CTA 0 of one tcgen05 GEMM launch (-DFHVAE_TIMELINE build).  Development aid."""
import ctypes, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_scalablefhvae_b200 import _lib
from pytorch_scalablefhvae_b200._lib import GemmProblem
from pytorch_scalablefhvae_b200.plan import ptr
so = "/tmp/libfhvae_tl.so"
srcs = [os.path.join(_lib.CSRC, s) for s in _lib.SOURCES]
subprocess.check_call(["nvcc"] + _lib.NVCC_FLAGS + ["-DFHVAE_TIMELINE", "-o", so] + srcs)
lib = ctypes.CDLL(so)
lib.fhvae_gemm_batch.argtypes = [ctypes.POINTER(GemmProblem), ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
for (M, N, K, kind) in [(5120, 1024, 256, "nt"), (128, 128, 256, "nt"), (1024, 256, 5120, "tn"), (256, 64, 512, "nt")]:
    if kind == "nt":
        A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.zeros(M, N, device="cuda")
        p = GemmProblem(ptr(A), ptr(B), ptr(C), None, M, N, K, 0, K, 1, 1, K, N, 0.0, 0)
    else:
        A = torch.randn(K, M, device="cuda"); B = torch.randn(K, N, device="cuda"); C = torch.zeros(M, N, device="cuda")
        p = GemmProblem(ptr(A), ptr(B), ptr(C), None, M, N, K, 0, 1, M, N, 1, N, 0.0, 0)
    arr = (GemmProblem * 1)(p)
    for _ in range(3):
        lib.fhvae_gemm_batch(arr, 1, 1, None)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 64)()
    lib.fhvae_debug_gemm_timeline(buf)
    t0 = buf[0]
    print(f"{kind} {M}x{N}x{K}: prologue {buf[1]-t0}; producer stage-done at", [buf[2+i]-t0 for i in range(8)],
          "; acc done", buf[20]-t0, "; epilogue done", buf[21]-t0)
    print("   MMA warp: full-wait done / commit issued at", [(buf[32+2*i]-t0, buf[33+2*i]-t0) for i in range(8)])
