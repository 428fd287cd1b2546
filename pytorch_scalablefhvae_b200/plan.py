"""Call lists: a train/inference step is a fixed sequence of C-ABI calls on preallocated buffers.

A ``CallList`` is built once per (model, batch size); replaying it costs one ctypes call per kernel
launch with pre-marshalled arguments, and is CUDA-graph capturable (the library never synchronises).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib
from ._lib import ColsumProblem, GemmProblem, ProjProblem


def ptr(t: torch.Tensor, off: int = 0) -> int:
    """Device address of element ``off`` (in elements) of a float32/int64 tensor's storage view."""
    return t.data_ptr() + off * t.element_size()


def gemm_nt(A, lda, W, ldw, Cp, ldc, M, N, K, bias=0, beta=0.0, relu=0) -> GemmProblem:
    """C[M,N] = A[M,K] @ W[N,K]^T (+bias) -- nn.Linear forward."""
    return GemmProblem(A, W, Cp, bias or None, M, N, K, relu, lda, 1, 1, ldw, ldc, beta, 0)


def gemm_nn(A, lda, W, ldw, Cp, ldc, M, N, K, beta=0.0) -> GemmProblem:
    """C[M,N] = A[M,K] @ W[K,N] -- data gradient dX = dY @ W."""
    return GemmProblem(A, W, Cp, None, M, N, K, 0, lda, 1, ldw, 1, ldc, beta, 0)


def gemm_tn(G, ldg, X, ldx, Cp, ldc, M, N, K, beta=0.0) -> GemmProblem:
    """C[M,N] = G[K,M]^T @ X[K,N] -- weight gradient dW = dY^T @ X (contraction over rows)."""
    return GemmProblem(G, X, Cp, None, M, N, K, 0, 1, ldg, ldx, 1, ldc, beta, 0)


_SIDE_STREAMS = {}


def _side_stream(device, k: int = 1) -> "torch.cuda.Stream":
    key = (torch.cuda.current_device() if device is None else device, k)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream()
    return _SIDE_STREAMS[key]


class CallList:
    """Ordered kernel launches.  A call tagged ``side=k`` (k = 1, 2, 3; True == 1) is off the critical path: it is
    forked onto side stream k behind an event recorded at its position in the main sequence (calls on one side
    stream stay in order), and joined back by ``join(k)`` or at the end of the list.  Stream 1 carries the weight
    gradients (they fill the SMs the 128-CTA recurrence leaves idle), stream 2 the short discriminative chain,
    stream 3 the bias column sums.  Works identically eagerly and under CUDA-graph capture (fork/join become graph branches)."""

    def __init__(self):
        self.calls = []      # (cfunc, name, args, side)
        self.keep = []       # ctypes arrays / tensors that must outlive the list

    def add(self, name: str, *args, side=False):
        self.calls.append((_lib.fn(name), name, args, int(side)))

    def gemm(self, problems: List[GemmProblem], mode: int, side=False):
        for i in range(0, len(problems), _lib.GEMM_MAX_BATCH):
            chunk = problems[i:i + _lib.GEMM_MAX_BATCH]
            arr = (GemmProblem * len(chunk))(*chunk)
            self.keep.append(arr)
            self.add("fhvae_gemm_batch", arr, len(chunk), mode, side=side)

    def proj(self, problems: List[ProjProblem], mode: int, side=False):
        """Projection GEMMs on bf16 hi/lo planes (csrc/gemm_proj.cu), grouped into one launch."""
        for i in range(0, len(problems), _lib.PROJ_MAX_BATCH):
            chunk = problems[i:i + _lib.PROJ_MAX_BATCH]
            arr = (ProjProblem * len(chunk))(*chunk)
            self.keep.append(arr)
            self.add("fhvae_proj_planes_batch", arr, len(chunk), mode, side=side)

    def colsum(self, problems: List[ColsumProblem], side=False):
        for i in range(0, len(problems), _lib.COLSUM_MAX_BATCH):
            chunk = problems[i:i + _lib.COLSUM_MAX_BATCH]
            arr = (ColsumProblem * len(chunk))(*chunk)
            self.keep.append(arr)
            self.add("fhvae_colsum_batch", arr, len(chunk), side=side)

    def torch_op(self, f):
        """A host callable (tiny torch ops on static buffers; still graph-capturable)."""
        self.calls.append((None, "torch", f, 0))

    def join(self, k: int = 1):
        """The main sequence waits here for everything issued so far on side stream k."""
        self.calls.append((None, "join", k, 0))

    def run(self, stream: int = 0, overlap: bool = True, join: bool = True):
        main = torch.cuda.current_stream()
        mptr = main.cuda_stream
        used = {}

        def join_side(k):
            s = used.pop(k, None)
            if s is not None:
                ev = torch.cuda.Event()
                ev.record(s)
                main.wait_event(ev)

        for f, name, args, on_side in self.calls:
            if f is None:
                if name == "join":
                    join_side(args)
                else:
                    args()
                continue
            if on_side and overlap:
                side = used.get(on_side)
                if side is None:
                    side = used[on_side] = _side_stream(None, on_side)
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                st = f(*args, side.cuda_stream)
            else:
                st = f(*args, mptr)
            if st:
                _lib.check(st, name)
        if join:
            for k in list(used):
                join_side(k)

    def __len__(self):
        return len(self.calls)


def current_stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream
