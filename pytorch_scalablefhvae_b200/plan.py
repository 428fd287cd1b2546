"""Call lists: a train/inference step is a fixed sequence of C-ABI calls on preallocated buffers.

A ``CallList`` is built once per (model, batch size); replaying it costs one ctypes call per kernel
launch with pre-marshalled arguments, and is CUDA-graph capturable (the library never synchronises).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib
from ._lib import ColsumProblem, GemmProblem


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


def _side_stream(device) -> "torch.cuda.Stream":
    key = torch.cuda.current_device() if device is None else device
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream()
    return _SIDE_STREAMS[key]


class CallList:
    """Ordered kernel launches.  Calls tagged ``side=True`` are off the critical path (weight / bias
    gradients): they are forked onto a second stream behind an event recorded at their position in the
    main sequence and joined back at the end, so they fill the SMs the 64-CTA cluster recurrence leaves
    idle.  Works identically eagerly and under CUDA-graph capture (fork/join become graph branches)."""

    def __init__(self):
        self.calls = []      # (cfunc, name, args, side)
        self.keep = []       # ctypes arrays / tensors that must outlive the list

    def add(self, name: str, *args, side: bool = False):
        self.calls.append((_lib.fn(name), name, args, side))

    def gemm(self, problems: List[GemmProblem], mode: int, side: bool = False):
        for i in range(0, len(problems), _lib.GEMM_MAX_BATCH):
            chunk = problems[i:i + _lib.GEMM_MAX_BATCH]
            arr = (GemmProblem * len(chunk))(*chunk)
            self.keep.append(arr)
            self.add("fhvae_gemm_batch", arr, len(chunk), mode, side=side)

    def colsum(self, problems: List[ColsumProblem], side: bool = False):
        for i in range(0, len(problems), _lib.COLSUM_MAX_BATCH):
            chunk = problems[i:i + _lib.COLSUM_MAX_BATCH]
            arr = (ColsumProblem * len(chunk))(*chunk)
            self.keep.append(arr)
            self.add("fhvae_colsum_batch", arr, len(chunk), side=side)

    def torch_op(self, f):
        """A host callable (tiny torch ops on static buffers; still graph-capturable)."""
        self.calls.append((None, "torch", f, False))

    def run(self, stream: int = 0, overlap: bool = True):
        main = torch.cuda.current_stream()
        mptr = main.cuda_stream
        side = None
        for f, name, args, on_side in self.calls:
            if f is None:
                args()
                continue
            if on_side and overlap:
                if side is None:
                    side = _side_stream(None)
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                st = f(*args, side.cuda_stream)
            else:
                st = f(*args, mptr)
            if st:
                _lib.check(st, name)
        if side is not None:
            ev = torch.cuda.Event()
            ev.record(side)
            main.wait_event(ev)

    def __len__(self):
        return len(self.calls)


def current_stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream
