"""ctypes binding of ``libfhvae_b200.so`` (the C ABI declared in ``include/fhvae_b200.h``).

There is no CPU fallback: if the shared library is missing, or a call returns non-zero, this
module raises.  ``build()`` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libfhvae_b200.so")
if os.environ.get("FHVAE_B200_LIB"):          # A/B experiments: a variant build of the same sources (tools/build_variant.sh)
    LIB_PATH = os.path.abspath(os.environ["FHVAE_B200_LIB"])
SOURCES = ["api.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm_wgrad.cu", "gemm_proj.cu", "lstm_simt.cu", "lstm_cluster.cu", "lstm_wave.cu", "elbo.cu", "heads.cu", "disc.cu",
           "table_adam_misc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

MODE_F32_SIMT, MODE_BF16X3, MODE_BF16 = 0, 1, 2
FLAG_NAN, FLAG_BAD_INDEX = 1, 2          # bits of the device status word (include/fhvae_b200.h)
GEMM_MAX_BATCH = 24
WGRAD_MAX_BATCH = 8
PROJ_MAX_BATCH = 8
SPLIT_MAX_BATCH = 16
COLSUM_MAX_BATCH = 16


class GemmProblem(C.Structure):
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("bias", C.c_void_p),
                ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("relu", C.c_int32),
                ("sa_m", C.c_int64), ("sa_k", C.c_int64), ("sb_k", C.c_int64), ("sb_n", C.c_int64),
                ("ldc", C.c_int64), ("beta", C.c_float), ("reserved", C.c_int32)]


class WgradProblem(C.Structure):
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("M", C.c_int32), ("N", C.c_int32),
                ("K", C.c_int32), ("reserved", C.c_int32), ("lda", C.c_int64), ("a_plane_stride", C.c_int64),
                ("ldb", C.c_int64), ("b_plane_stride", C.c_int64), ("ldc", C.c_int64)]


class ProjProblem(C.Structure):
    _fields_ = [("A", C.c_void_p), ("W", C.c_void_p), ("C", C.c_void_p), ("bias", C.c_void_p), ("M", C.c_int32),
                ("N", C.c_int32), ("K", C.c_int32), ("reserved", C.c_int32), ("lda", C.c_int64),
                ("a_plane_stride", C.c_int64), ("ldw", C.c_int64), ("w_plane_stride", C.c_int64), ("ldc", C.c_int64)]


class SplitProblem(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("ld_src", C.c_int64), ("ld_dst", C.c_int64),
                ("plane_stride", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32)]


class ColsumProblem(C.Structure):
    _fields_ = [("inp", C.c_void_p), ("out", C.c_void_p), ("out2", C.c_void_p), ("ld", C.c_int64),
                ("R", C.c_int32), ("C", C.c_int32)]


_p, _i, _l, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes (all return int except the two noted below); mirrors include/fhvae_b200.h
PROTOTYPES = {
    "fhvae_gemm_batch": [C.POINTER(GemmProblem), _i, _i, _p],
    "fhvae_lstm_fwd": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "fhvae_lstm_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "fhvae_lstm_wave_supported": [_i, _i, _i, _i, _i],
    "fhvae_lstm_wave_xchg_bytes": [_i, _i, _i, _i],
    "fhvae_lstm_wave_rows_per_launch": [_i, _i],
    "fhvae_lstm_wave_fwd": [_p] * 13 + [_i, _i, _i, _i, _i, _p],
    "fhvae_lstm_wave_bwd_xchg_bytes": [_i, _i, _i, _i],
    "fhvae_lstm_wave_fwd_planes": [_p] * 15 + [_l, _p, _i, _i, _i, _i, _i, _p],
    "fhvae_lstm_wave_bwd_planes": [_p] * 17 + [_l, _p, _i, _i, _i, _i, _i, _p],
    "fhvae_lstm_wave_pack_bytes": [_i, _i, _i],
    "fhvae_lstm_wave_pack": [_p, _p, _p, _p, _i, _i, _i, _p],
    "fhvae_lstm_wave_bwd": [_p] * 15 + [_i, _i, _i, _i, _i, _p],
    "fhvae_reparam_fwd": [_p, _l, _p, _p, _l, _i, _i, _p],
    "fhvae_reparam_bwd": [_p, _l, _p, _p, _l, _p, _l, _i, _i, _i, _p],
    "fhvae_elbo_fwd": [_p, _p, _l, _l, _l, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "fhvae_elbo_bwd": [_p, _p, _l, _l, _l, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "fhvae_elbo_fwd_bwd": [_p, _p, _l, _l, _l, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "fhvae_disc_nsplit": [_i, _l],
    "fhvae_disc_fwd_partial": [_p, _l, _p, _l, _i, _p, _i, _i, _p],
    "fhvae_disc_target": [_p, _l, _p, _p, _i, _i, _p],
    "fhvae_disc_combine": [_p, _i, _p, _p, _p, _i, _p],
    "fhvae_disc_bwd_rows": [_p, _l, _p, _l, _i, _p, _p, _p, _i, _p],
    "fhvae_disc_bwd_segs": [_p, _l, _p, _l, _i, _p, _p, _i, _i, _p],
    "fhvae_disc_bwd_finish": [_p, _l, _p, _p, _i, _p, _p, _l, _p, _i, _i, _p],
    "fhvae_shard_pack": [_p, _l, _p, _p, _p, _i, _i, _p],
    "fhvae_shard_unpack": [_p, _i, _i, _i, _i, _l, _p, _p, _p, _p, _p],
    "fhvae_disc_combine_sharded": [_p, _i, _i, _p, _i, _i, _p, _p, _p],
    "fhvae_mu2_gather": [_p, _p, _p, _i, _i, _l, _p, _p],
    "fhvae_mu2_scatter_reduce": [_p, _p, _p, _p, _i, _i, _l, _p],
    "fhvae_mu2_accumulate": [_p, _l, _p, _p, _p, _i, _i, _l, _p, _p],
    "fhvae_mu2_estimate_finish": [_p, _p, _p, _f, _l, _i, _p],
    "fhvae_rows_copy": [_p, _p, _p, _p, _l, _i, _p],
    "fhvae_adam_flat": [_p, _p, _p, _p, _l, _f, _f, _f, _f, _f, _p, _p, _p],
    "fhvae_randn": [_p, _l, C.c_uint64, _p, _p, _p],
    "fhvae_transpose_bt": [_p, _p, _i, _i, _i, _p],
    "fhvae_gather_segments": [_p, _p, _p, _p, _p, _i, _i, _i, _l, _p],
    "fhvae_colsum_batch": [C.POINTER(ColsumProblem), _i, _p],
    "fhvae_add2": [_p, _p, _p, _l, _p],
    "fhvae_relu_bwd": [_p, _p, _l, _p],
    "fhvae_axpy": [_p, _p, _f, _l, _p],
    "fhvae_wgrad_planes_batch": [C.POINTER(WgradProblem), _i, _i, _p],
    "fhvae_split_planes_batch": [C.POINTER(SplitProblem), _i, _p],
    "fhvae_proj_planes_batch": [C.POINTER(ProjProblem), _i, _i, _p],
    "fhvae_load_inputs": [_p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p],
    "fhvae_head_fwd": [_p, _p, _l, _i, _i, _p, _p, _p, _i, _p, _p, _l, _i, _p, _l, _p, _i, _i, _p, _i, _i, _p],
    "fhvae_head_bwd": [_p, _i, _p, _l, _i, _p, _l, _i, _i, _p, _p, _i, _i, _p, _i, _p, _i, _i, _p, _p, _i, _p],
    "fhvae_step_coef": [_p, _p, _p, _i, _i, _i, _p],
    "fhvae_loss_mean": [_p, _p, _f, _i, _p, _p],
    "fhvae_set_deterministic": [_i],
    "fhvae_get_deterministic": [],
    "fhvae_version": [],
    "fhvae_built_for_sm": [],
    "fhvae_launch_count": [],
}
NO_STATUS = {"fhvae_lstm_wave_rows_per_launch", "fhvae_set_deterministic", "fhvae_get_deterministic", "fhvae_lstm_wave_pack_bytes", "fhvae_disc_nsplit", "fhvae_lstm_wave_supported", "fhvae_lstm_wave_xchg_bytes", "fhvae_lstm_wave_bwd_xchg_bytes", "fhvae_version", "fhvae_built_for_sm", "fhvae_launch_count"}
EXPORTS = sorted(list(PROTOTYPES) + ["fhvae_last_error_string"])

_lib = None


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile csrc/*.cu -> libfhvae_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "fhvae_b200.h"))
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps)):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built (run "
            "`python -c 'import __graft_entry__ as g; g.build()'`).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.fhvae_last_error_string.restype = C.c_char_p
    lib.fhvae_last_error_string.argtypes = []
    for name, args in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.argtypes = args
        fn.restype = (C.c_ulonglong if name == "fhvae_launch_count" else
                      C.c_longlong if name.endswith("_bytes") else C.c_int)
    _lib = lib
    return lib


class FhvaeError(RuntimeError):
    pass


def check(status: int, name: str = ""):
    if status != 0:
        msg = load().fhvae_last_error_string().decode()
        raise FhvaeError(f"{name or 'libfhvae_b200'} failed with status {status}: {msg}")


def set_deterministic(on: bool = True) -> bool:
    """Fixed summation order in every split-K launch (two bit-identical runs of a step); returns the previous
    setting.  Call before the first step: captured CUDA graphs bake the launch geometry in."""
    return bool(load().fhvae_set_deterministic(1 if on else 0))


def fn(name: str):
    """Return the raw ctypes function (status must be checked by the caller)."""
    return getattr(load(), name)
