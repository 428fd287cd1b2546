"""``SimpleFHVAE`` / ``FHVAE``: drop-in ``nn.Module`` surface of the reference, computed by
hand-written sm_100a kernels behind the C ABI of ``include/fhvae_b200.h``.

Mirrors (file:line into BurnhamG/PyTorch-ScalableFHVAE):
* constructors            simple_fhvae.py:9-37, fhvae.py:5-13 (``FHVAE`` is a stub upstream: the LSTM
                          architecture is SURVEY.md Appendix B, frozen in oracle/fhvae_oracle.py)
* ``forward`` 4-arg call  simple_fhvae.py:71-124, called from train_model.py:447-449 / utils.py:51
* attribute surface       ``.model .z1_hus .z2_hus .z1_dim .z2_dim .x_hus .pz1 .pmu2`` (utils.py:72-77,
                          134-141) plus ``.qz2_x .qz1_x .px_z .pz2`` which utils.estimate_mu2_dict reads
                          (utils.py:52,58) but the reference forgets to set.

There is no CPU path: parameters live in one flat fp32 HBM buffer (``nn.Parameter`` views of it, so
``named_parameters()`` / ``state_dict()`` / ``.grad`` behave as usual), a forward/backward is a
replayed list of kernel launches on preallocated workspaces, optionally captured in a CUDA graph.
"""
from __future__ import annotations

import math
import os
import weakref
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import ColsumProblem, ProjProblem, SplitProblem, WgradProblem
from .plan import CallList, _side_stream, current_stream_ptr, gemm_nn, gemm_nt, gemm_tn, ptr

PZ2_LOGVAR = math.log(0.5 ** 2)      # simple_fhvae.py:88
PMU2_LOGVAR = math.log(1.0 ** 2)     # simple_fhvae.py:23
_OUT_ORDER = (0, 5, 1, 2, 3, 4)      # rows of the (6,B) out buffer -> reference return order


class _Holder(nn.Module):
    """Name-space node so that parameter keys match the reference's state_dict."""


def _prod(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


def _as_int_list(v) -> List[int]:
    # train_model.py:145-168 forgets type=int for --z1-hus/--z2-hus/--x-hus: strings may arrive
    return [int(h) for h in v]


class _FHVAECore(nn.Module):
    """Flat parameter storage, plan cache, autograd glue and the shared forward tail."""

    model = "base"
    # How the backward hands the gradients to the parameters.  True (default): the step node assigns cached views of
    # the flat gradient buffer to ``p.grad`` itself (what ``loss.backward()`` callers, optimizers, clipping and the
    # reference's TensorBoard histograms see is identical; the per-step host cost of ~45 parameter edges through
    # the autograd engine -- ~0.5 ms, more than the GPU time it would hide -- disappears).  False: the parameters
    # are inputs of the node and autograd's AccumulateGrad delivers the same views (needed only for
    # ``torch.autograd.grad`` w.r.t. parameters or parameter hooks).
    direct_grads = True

    # ------------------------------------------------------------------ parameters
    def _init_flat(self, specs: Sequence[Tuple[str, Tuple[int, ...]]], init: Dict[str, torch.Tensor]):
        off = 0
        self._off: Dict[str, int] = {}
        self._shape: Dict[str, Tuple[int, ...]] = {}
        for name, shape in specs:
            self._off[name] = off
            self._shape[name] = tuple(shape)
            off += (_prod(shape) + 3) // 4 * 4          # 16-byte aligned slots
        self._n_flat = off
        # registration order = the reference's construction order (the insertion order of `init`), so that
        # named_parameters() / optimizer.state_dict() indices line up with the reference's modules; the
        # flat-buffer LAYOUT (specs order) is independent of it
        order = {n: i for i, n in enumerate(init.keys())}
        specs = sorted(specs, key=lambda ns: order[ns[0]])
        self._names = [n for n, _ in specs]
        flat = torch.zeros(off, dtype=torch.float32)
        self._plist: List[nn.Parameter] = []
        for name, shape in specs:
            view = flat[self._off[name]:self._off[name] + _prod(shape)].view(shape)
            view.copy_(init[name].detach().float())
            p = nn.Parameter(view)
            p._fhvae = (weakref.ref(self), name)
            node = self
            parts = name.split(".")
            for part in parts[:-1]:
                if not hasattr(node, part):
                    node.add_module(part, _Holder())
                node = getattr(node, part)
            node.register_parameter(parts[-1], p)
            self._plist.append(p)
        self._flat = flat
        self._gflat: List[Optional[torch.Tensor]] = [None, None]
        self._plans: Dict[Tuple, "_Plan"] = {}
        self._anchor = None

    def _ensure_flat(self) -> torch.Tensor:
        """(Re)pack parameters into one flat buffer after .to()/.cuda()/load; keeps Parameter identity."""
        p0, pl = self._plist[0], self._plist[-1]
        ok = (self._flat.device == p0.device
              and p0.data_ptr() == ptr(self._flat, self._off[self._names[0]])
              and pl.data_ptr() == ptr(self._flat, self._off[self._names[-1]]))
        if ok:
            return self._flat
        dev = p0.device
        flat = torch.zeros(self._n_flat, dtype=torch.float32, device=dev)
        for name, p in zip(self._names, self._plist):
            if p.dtype != torch.float32:
                raise RuntimeError("the sm_100a path is fp32 only (SURVEY.md §8b dtype contract)")
            o, n = self._off[name], p.numel()
            view = flat[o:o + n].view(self._shape[name])
            view.copy_(p.data)
            p.data = view
        self._flat = flat
        self._gflat = [None, None]
        self._plans.clear()
        return flat

    def double(self):          # train_model.py:438 -- unsupported on the CUDA path (Appendix A8)
        raise RuntimeError(
            "FHVAE sm_100a kernels are fp32; model.double() (train_model.py:438) is not supported. "
            "fp32 matches the fp64 reference to ~1e-6 relative (SURVEY.md Appendix D).")

    def poff(self, name: str, extra: int = 0) -> int:
        return ptr(self._flat, self._off[name] + extra)

    def _grad_buffer(self, k: int) -> torch.Tensor:
        if self._gflat[k] is None or self._gflat[k].device != self._flat.device:
            self._gflat[k] = torch.zeros(self._n_flat, dtype=torch.float32, device=self._flat.device)
        return self._gflat[k]

    def _free_grad_slot(self) -> int:
        """Index of a flat grad buffer that no live ``.grad`` aliases (autograd glue)."""
        g = self._plist[0].grad
        if g is None:
            return 0
        for k in (0, 1):
            buf = self._gflat[k]
            if buf is None or g.data_ptr() != ptr(buf, self._off[self._names[0]]):
                return k
        return 0

    def _grad_views_of(self, k: int) -> List[torch.Tensor]:
        """Per-parameter views of flat gradient buffer k (created once per buffer)."""
        buf = self._grad_buffer(k)
        cache = self.__dict__.setdefault("_gviews", {})
        ent = cache.get(k)
        if ent is None or ent[0] is not buf:
            ent = (buf, [buf[self._off[n]:self._off[n] + p.numel()].view(self._shape[n])
                         for n, p in zip(self._names, self._plist)])
            cache[k] = ent
        return ent[1]

    def _deliver_grads(self, k: int):
        """``p.grad`` <- view of buffer k (or accumulate into an existing ``p.grad``, like AccumulateGrad)."""
        for p, v in zip(self._plist, self._grad_views_of(k)):
            if not p.requires_grad:
                continue
            if p.grad is None:
                p.grad = v
            else:
                p.grad.add_(v)

    def packed_grads(self) -> torch.Tensor:
        """Flat gradient buffer equal to every ``p.grad`` (zeros where None); zero-copy on the fast path."""
        self._ensure_flat()
        for k, ent in self.__dict__.get("_gviews", {}).items():     # identity check against the cached views
            if ent[0] is self._gflat[k] and all(p.grad is v for p, v in zip(self._plist, ent[1])):
                return ent[0]
        for k in (0, 1):
            buf = self._gflat[k]
            if buf is None:
                continue
            if all(p.grad is not None and p.grad.data_ptr() == ptr(buf, self._off[n])
                   for n, p in zip(self._names, self._plist)):
                return buf
        k = self._free_grad_slot()
        buf = self._grad_buffer(k)
        buf.zero_()
        for n, p in zip(self._names, self._plist):
            if p.grad is not None:
                buf[self._off[n]:self._off[n] + p.numel()].view_as(p).copy_(p.grad)
        return buf

    # ------------------------------------------------------------------ forward
    def _padded_batch(self, B: int, T: int) -> int:
        """Rows the plan of a B-segment batch is built for.  The tensor-core recurrence works on groups of 32 batch rows; a
        ragged batch (the last one of an epoch: the reference's DataLoader keeps it, train_model.py:440) is run on the next
        multiple of 32 with the extra rows carrying finite filler and ZERO upstream gradient, instead of falling back to
        the per-step fp32 kernels (~10x slower).  Sub-classes without such kernels return B."""
        return B

    def _plan(self, B: int, T: int, F: int) -> "_Plan":
        self._ensure_flat()
        key = (B, T, F, self.gemm_mode, self._flat.data_ptr(), _lib.load().fhvae_get_deterministic())
        plan = self._plans.get(key)
        if plan is None:
            plan = self._make_plan(B, T, F)
            self._plans[key] = plan
        return plan

    def forward(self, x: torch.Tensor, mu_idx: torch.Tensor, num_seqs: int, num_segs, eps=None):
        """simple_fhvae.py:71-124.  Returns (lower_bound (B,), log_qy, log_px_z (B,), neg_kld_z1 (B,),
        neg_kld_z2 (B,), log_pmu2 (B,)); ``log_qy`` is per-segment -CE_b unless ``ref_log_qy``."""
        if not x.is_cuda:
            raise RuntimeError("pytorch_scalablefhvae_b200 has no CPU path: move the batch and the "
                               "model to a CUDA device (train_model.py:413,444)")
        if self._plist[0].device != x.device:
            raise RuntimeError(f"model parameters are on {self._plist[0].device}, input on {x.device}: "
                               "call model.to(device) first (train_model.py:413)")
        if int(num_seqs) != self.mu2_table.shape[0]:
            raise ValueError(f"num_seqs={num_seqs} but the mu2 table has {self.mu2_table.shape[0]} rows")
        B, T, F = x.shape
        self._check_ids(mu_idx, num_segs, B, num_seqs)
        plan = self._plan(self._padded_batch(B, T), T, F)
        plan.load_inputs(x, mu_idx, num_segs, eps, valid=B)
        grad_on = torch.is_grad_enabled() and any(p.requires_grad for p in self._plist)
        if grad_on:
            if self._anchor is None or self._anchor.device != x.device:
                self._anchor = torch.zeros(1, device=x.device, requires_grad=True)
            params = () if self.direct_grads else self._plist
            out = _StepFn.apply(self, plan, self._anchor, *params)          # six (B,) rows of one buffer
        else:
            plan.run_forward()
            out = plan.out[:, :B].clone().unbind(0)
        try:
            self._check_id_range(mu_idx, B, num_seqs)
        except IndexError:
            plan.nan_flag.zero_()        # the launches above flagged the same ids on the device: reported here, once
            raise
        self._publish(plan)
        lb, log_qy, log_px_z, nk1, nk2, log_pmu2 = (out[i] for i in _OUT_ORDER)
        if self.ref_log_qy:                      # reference returns mean(+CE) (simple_fhvae.py:37,122)
            log_qy = -log_qy.mean()
        return lb, log_qy, log_px_z, nk1, nk2, log_pmu2

    def train_step(self, x, mu_idx, num_segs, optimizer, alpha: float = 10.0, eps=None, allreduce=None, shard=None,
                   overlap=None):
        """Fused loop body of train_model.py:446-454: forward, loss = -mean(lb + alpha*log_qy)
        (train_model.py:243-251), backward, Adam -- one replayed launch sequence (one CUDA graph when
        ``use_cuda_graphs`` and no collective is interposed).  ``allreduce(flat_grads)`` is called
        between backward and Adam for data-parallel training.  Returns the loss (device scalar)."""
        if not x.is_cuda and not x.is_pinned():
            raise RuntimeError("train_step takes a CUDA or pinned-host batch (no CPU path)")
        B, T, F = x.shape
        # with a sharded table (parallel.DataParallel(table="sharded")) mu_idx are GLOBAL row ids in [0, num_rows)
        n_rows = shard.num_rows if shard is not None else self.mu2_table.shape[0]
        self._check_ids(mu_idx, num_segs, B, n_rows)
        self._check_id_range(mu_idx, B, n_rows)      # before anything is queued: the fused step would apply Adam
        fused_dp = shard is not None or (overlap is not None and allreduce is not None)
        plan = self._plan(B if fused_dp else self._padded_batch(B, T), T, F)    # (the sharded / overlapped runners: no padding)
        plan.load_inputs(x, mu_idx, num_segs, eps, valid=B)
        if shard is not None:
            return plan.run_train_step_sharded(optimizer, float(alpha), shard)
        if overlap is not None and allreduce is not None and hasattr(self, "_z2_end") and self.use_cuda_graphs:
            return plan.run_train_step_overlapped(optimizer, float(alpha), overlap)
        return plan.run_train_step(optimizer, float(alpha), allreduce)

    @staticmethod
    def _check_ids(mu_idx, num_segs, B, N):
        """torch.gather's error behaviour (simple_fhvae.py:53).  Host-resident ids (the reference keeps them on the
        CPU, train_model.py:445) are range-checked here; device-resident ids cannot be checked without a sync: the
        kernels never touch memory outside the table for them, poison the segment's lower bound with NaN and set
        FHVAE_FLAG_BAD_INDEX in ``model.nan_flag`` (``check_flags()`` reads it)."""
        if mu_idx.numel() != B:
            raise IndexError(f"mu_idx has {mu_idx.numel()} entries for a batch of {B} segments")
        if torch.is_tensor(num_segs) and num_segs.numel() != B:
            raise IndexError(f"num_segs has {num_segs.numel()} entries for a batch of {B} segments")

    @staticmethod
    def _check_id_range(mu_idx, B, N):
        """Range check of host-resident ids.  forward() calls it AFTER the step's launches are queued (the host check then
        overlaps the GPU; the kernels themselves never index outside the table, see ``_check_ids``)."""
        if not mu_idx.is_cuda and B:
            lo, hi = torch.aminmax(mu_idx)
            if int(lo) < 0 or int(hi) >= int(N):
                raise IndexError("mu_idx out of range for the mu2 table")

    def check_flags(self):
        """Host read (one sync) of the device status word: raises like the reference would have
        (IndexError of torch.gather, simple_fhvae.py:53; NaN lower bound, train_model.py:464-466)."""
        flag = int(self.nan_flag) if "nan_flag" in self.__dict__ else 0
        for plan in self._plans.values():
            flag |= int(plan.nan_flag)
        if flag & _lib.FLAG_BAD_INDEX:
            raise IndexError("mu_idx out of range for the mu2 table (device-side check)")
        if flag & _lib.FLAG_NAN:
            raise FloatingPointError("NaN in the lower bound (train_model.py:464)")

    @torch.no_grad()
    def encode(self, x: torch.Tensor, eps=None):
        """Forward-only posterior extraction (what eval_model.py:55-59 leaves as TODO): runs only the two
        encoders.  With ``eps=None`` the z2 *mean* conditions the z1 encoder (eps = 0).  Returns views of
        the plan's buffers: dict(z1_mu, z1_logvar, z2_mu, z2_logvar), valid until the next call."""
        if not x.is_cuda:
            raise RuntimeError("pytorch_scalablefhvae_b200 has no CPU path")
        B, T, F = x.shape
        plan = self._plan(self._padded_batch(B, T), T, F)
        plan.set_x(x)
        if eps is None:
            plan.eps1.zero_(); plan.eps2.zero_()
        else:
            plan.eps1[:B].copy_(eps["z1"].reshape(B, -1)); plan.eps2[:B].copy_(eps["z2"].reshape(B, -1))
        plan.run_encode()
        Z1, Z2 = self.z1_dim, self.z2_dim
        return {"z1_mu": plan.z1head[:B, :Z1], "z1_logvar": plan.z1head[:B, Z1:],
                "z2_mu": plan.z2head[:B, :Z2], "z2_logvar": plan.z2head[:B, Z2:]}

    def _publish(self, plan: "_Plan"):
        """Attributes the reference's callers read (utils.py:52,58)."""
        pub = plan.__dict__.get("_published")
        v = plan.valid
        if pub is None or pub[0] != v:     # views of the plan's static buffers: built once, not per forward
            z1h, z2h = plan.z1head[:v], plan.z2head[:v]
            Z1, Z2 = self.z1_dim, self.z2_dim
            pub = plan._published = (v, dict(
                qz1_x=[z1h[:, :Z1], z1h[:, Z1:]], qz2_x=[z2h[:, :Z2], z2h[:, Z2:]], px_z=plan.px_views(),
                pz2=[plan.mu2[:v], np.float32(PZ2_LOGVAR)], z1_sample=plan.zcat[:v, :Z1], z2_sample=plan.zcat[:v, Z1:],
                nan_flag=plan.nan_flag))
        self.__dict__.update(pub[1])

    # subclasses: _make_plan(B, T, F)


class _StepFn(torch.autograd.Function):
    """One autograd node for the whole step: forward replays the forward call list, backward the
    BPTT/wgrad list, returning views of one flat gradient buffer (no per-parameter kernels)."""

    @staticmethod
    def forward(ctx, model, plan, anchor, *params):
        plan.run_forward()
        ctx.model, ctx.plan, ctx.n_params = model, plan, len(params)
        ctx.generation = plan.generation
        ctx.set_materialize_grads(False)      # unused outputs arrive as None instead of zero tensors
        # rows of the (6,B) buffer: lb, log_px, nk1, nk2, log_pmu2, log_qy -- returned as six outputs so that
        # autograd hands their gradients straight back (no per-row SelectBackward zeros + copy)
        return tuple(plan.out[:, :plan.valid].clone().unbind(0))

    @staticmethod
    def backward(ctx, *gouts):
        model, plan = ctx.model, ctx.plan
        if plan.generation != ctx.generation:
            # the backward replays on the plan's static activation buffers: a later forward / encode / train_step of
            # the same (B,T,F) has overwritten what this node saved (INTEGRATION.md "one step in flight")
            raise RuntimeError(
                "pytorch_scalablefhvae_b200: backward() of a forward whose activations were overwritten by a later "
                "forward/encode/train_step of the same batch shape; call backward() before the next forward "
                "(one step in flight per batch shape, INTEGRATION.md)")
        if getattr(ctx, "consumed", False):
            raise RuntimeError("pytorch_scalablefhvae_b200: double backward through one step is not supported")
        ctx.consumed = True
        k = model._free_grad_slot()
        plan.run_backward(gouts, k)
        if ctx.n_params == 0:                 # direct delivery (see _FHVAECore.direct_grads)
            model._deliver_grads(k)
            return (None, None, None)
        gflat = model._grad_buffer(k)         # fresh views: AccumulateGrad steals un-shared tensors (zero copies)
        grads = [gflat[model._off[n]:model._off[n] + p.numel()].view(model._shape[n])
                 if p.requires_grad else None
                 for n, p in zip(model._names, model._plist)]
        return (None, None, None, *grads)


class _LossFn(torch.autograd.Function):
    """-mean(lower_bound + alpha*log_qy) (train_model.py:243-251) as one launch; the backward is two scalings."""

    @staticmethod
    def forward(ctx, lb, log_qy, alpha):
        B = lb.numel()
        loss = torch.empty((), dtype=torch.float32, device=lb.device)
        _lib.check(_lib.fn("fhvae_loss_mean")(ptr(lb), ptr(log_qy), float(alpha), B, ptr(loss), current_stream_ptr()),
                   "fhvae_loss_mean")
        ctx.B, ctx.alpha = B, float(alpha)
        return loss

    @staticmethod
    def backward(ctx, g):
        glb = (g * (-1.0 / ctx.B)).expand(ctx.B)
        gqy = (g * (-ctx.alpha / ctx.B)).expand(ctx.B)
        return glb, gqy, None


# =====================================================================================
# plans
# =====================================================================================
class _Plan:
    """Workspaces + call lists for one (model, B, T, F).  Sub-classes fill fwd / bwd lists."""

    def __init__(self, m: _FHVAECore, B: int, T: int, F: int):
        self.m, self.B, self.T, self.F = m, B, T, F
        self.dev = m._flat.device
        self.mode = m.gemm_mode
        Z1, Z2, N = m.z1_dim, m.z2_dim, m.mu2_table.shape[0]
        self.Z1, self.Z2, self.N = Z1, Z2, N
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.dev)
        self.f = f
        # static inputs
        self.x = f(B, T, F)
        self.idx = torch.zeros(B, dtype=torch.int64, device=self.dev)
        self.nsegs = torch.ones(B, dtype=torch.int64, device=self.dev)
        self.eps_all = f(B * (Z1 + Z2))                # [eps2 | eps1]: one normal_() launch draws both
        self.eps2, self.eps1 = self.eps_all[:B * Z2].view(B, Z2), self.eps_all[B * Z2:].view(B, Z1)
        # eps ~ N(0,1) (torch.randn_like, simple_fhvae.py:214; z2 first, then z1) drawn by the library's own Philox kernel
        # as the first node of the step: seeded from torch's generator when the plan is built, advanced on the device
        self._rng_state = torch.zeros(2, dtype=torch.int64, device=self.dev)       # [offset, done counter]
        self._rng_seed = (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + B * 1000003 + T * 10007 + F) & (2 ** 64 - 1)
        self._draw = True
        self.rng = CallList()
        self.rng.add("fhvae_randn", ptr(self.eps_all), self.eps_all.numel(), self._rng_seed, ptr(self._rng_state),
                     ptr(self._rng_state, 1))
        # forward state
        self.z1head, self.z2head = f(B, 2 * Z1), f(B, 2 * Z2)
        self.zcat = f(B, Z1 + Z2)                     # [z1_sample | z2_sample]
        self.mu2 = f(B, Z2)
        self.out = f(6, B)                            # lb, log_px, nk1, nk2, log_pmu2, log_qy
        self.nsplit = _lib.fn("fhvae_disc_nsplit")(B, N)
        self.part = f(self.nsplit, B, 2)
        self.tgt, self.lse = f(B), f(B)
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.fwd = CallList()
        self.bwd: List[Optional[CallList]] = [None, None]
        self.gout = f(6, B)
        self._graph_fwd = None
        self._graph_bwd = [None, None]
        self.generation = 0          # bumped by everything that overwrites the saved activations (see _StepFn.backward)
        self.valid = B               # rows of the current batch (< B for a ragged batch run on a padded plan)

    # ---- inputs / outputs
    def set_x(self, x, mu_idx=None, num_segs=None):
        """x -> the plan's static (B,T,F) buffer and, where the plan keeps one, its time-major copy; optionally the
        two id vectors in the same launch.  Every entry (forward, train_step, encode) loads x through here."""
        self.generation += 1
        x_tm = self.__dict__.get("x_tm")
        self.valid = int(x.shape[0])
        if self.valid < self.B:      # ragged batch on a padded plan: the filler rows keep whatever finite data they hold
            self.x[:self.valid].copy_(x, non_blocking=True)
            if x_tm is not None:
                _lib.check(_lib.fn("fhvae_transpose_bt")(ptr(self.x), ptr(x_tm), self.B, self.T, self.F,
                                                         current_stream_ptr()), "fhvae_transpose_bt")
            return
        if (x.device == self.dev and x.dtype == torch.float32 and x.is_contiguous() and x.shape == self.x.shape
                and self.F % 4 == 0 and x.data_ptr() % 16 == 0):
            _lib.check(_lib.fn("fhvae_load_inputs")(
                ptr(x), ptr(self.x), ptr(x_tm) if x_tm is not None else None, self.B, self.T, self.F,
                ptr(mu_idx) if mu_idx is not None else None, ptr(self.idx),
                ptr(num_segs) if num_segs is not None else None, ptr(self.nsegs), current_stream_ptr()),
                "fhvae_load_inputs")
            return
        self.x.copy_(x, non_blocking=True)
        if x_tm is not None:
            _lib.check(_lib.fn("fhvae_transpose_bt")(ptr(self.x), ptr(x_tm), self.B, self.T, self.F,
                                                     current_stream_ptr()), "fhvae_transpose_bt")
        if mu_idx is not None:
            self.idx.copy_(mu_idx, non_blocking=True)
            self.nsegs.copy_(num_segs, non_blocking=True)

    def load_inputs(self, x, mu_idx, num_segs, eps, valid=None):
        dev = self.dev
        v = self.B if valid is None else int(valid)
        # ids the load kernel can read itself: device-resident, or PINNED host tensors (mapped into the device's address
        # space: 4 KB over PCIe inside the launch instead of two cudaMemcpyAsync calls, ~15 us of host time per step)
        on_dev = lambda t: (torch.is_tensor(t) and t.dtype == torch.int64 and t.is_contiguous() and t.numel() == self.B
                            and (t.device == dev or (t.device.type == "cpu" and t.is_pinned())))
        fused_ids = v == self.B and on_dev(mu_idx) and on_dev(num_segs)
        self.set_x(x, mu_idx if fused_ids else None, num_segs if fused_ids else None)
        if not fused_ids:
            self.idx[:v].copy_(mu_idx, non_blocking=True)
            if torch.is_tensor(num_segs):
                self.nsegs[:v].copy_(num_segs, non_blocking=True)
            else:
                self.nsegs.fill_(int(num_segs))
        if eps is not None:
            self.eps1[:v].copy_(eps["z1"].reshape(v, -1), non_blocking=True)
            self.eps2[:v].copy_(eps["z2"].reshape(v, -1), non_blocking=True)
        self._draw = eps is None                     # True: the step itself draws eps (self.rng, first node of the graph)

    # ---- execution
    def run_forward(self):
        draw = self._draw

        def body():
            if draw:
                self.rng.run(current_stream_ptr())
            self.fwd.run(current_stream_ptr())
        if self.m.use_cuda_graphs:
            if self._graph_fwd is None:
                self._graph_fwd = {}
            if draw not in self._graph_fwd:
                self._graph_fwd[draw] = self._capture(body)
            self._graph_fwd[draw].replay()
        else:
            body()

    def run_encode(self):
        """Replay only the encoder prefix of the forward list (marked by ``n_encode_calls``)."""
        enc = self.__dict__.get("_enc")
        if enc is None:
            enc = CallList()
            enc.calls = self.fwd.calls[:self.n_encode_calls]
            enc.keep = self.fwd.keep
            self._enc = enc
        if self.m.use_cuda_graphs:
            if self.__dict__.get("_graph_enc") is None:
                self._graph_enc = self._capture(lambda: enc.run(current_stream_ptr()))
            self._graph_enc.replay()
        else:
            enc.run(current_stream_ptr())

    def run_backward(self, gout, k: int) -> torch.Tensor:
        gflat = self.m._grad_buffer(k)
        if self.bwd[k] is None:
            self.bwd[k] = self._build_bwd(gflat)
        v = self.valid
        if torch.is_tensor(gout):
            if v < self.B:
                self.gout.zero_()
            self.gout[:, :v].copy_(gout)
            self._gout_rows = None
        else:                                  # six per-output gradients, None where an output was not used
            used = (v,) + tuple(g_ is not None for g_ in gout)
            if self.__dict__.get("_gout_rows") != used:      # rows that are None now may hold an older gradient; the
                self.gout.zero_()                            # filler rows of a ragged batch must stay zero
                self._gout_rows = used
            for row, g_ in enumerate(gout):
                if g_ is not None:
                    self.gout[row, :v].copy_(g_)
        self._gout_train = None
        if self.m.use_cuda_graphs:
            if self._graph_bwd[k] is None:
                self._graph_bwd[k] = self._capture(lambda: self.bwd[k].run(current_stream_ptr()))
            self._graph_bwd[k].replay()
        else:
            self.bwd[k].run(current_stream_ptr())
        return gflat

    def run_train_step(self, optimizer, alpha, allreduce):
        m = self.m
        k = 0
        gflat = m._grad_buffer(k)
        if self.bwd[k] is None:
            self.bwd[k] = self._build_bwd(gflat)
        if not hasattr(self, "loss"):
            self.loss = torch.zeros((), dtype=torch.float32, device=self.dev)
        B, v = self.B, self.valid

        # the upstream gradients of a plain train step are constants: filled once, outside the graph
        # (filler rows of a ragged batch: zero, so that nothing they compute reaches a gradient)
        if self.__dict__.get("_gout_train") != (alpha, v):
            self.gout.zero_()
            self.gout[0, :v].fill_(-1.0 / v)              # d loss / d lower_bound
            self.gout[5, :v].fill_(-alpha / v)            # d loss / d log_qy
            self._gout_train = (alpha, v)
            self._gout_rows = None
        if self.__dict__.get("_loss_call") is None or self._loss_call[1] != (alpha, v):
            lc = CallList()
            lc.add("fhvae_loss_mean", ptr(self.out), ptr(self.out, 5 * B), float(alpha), v, ptr(self.loss), side=1)
            self._loss_call = (lc, (alpha, v))
        loss_call = self._loss_call[0]

        fwd_list, bwd_list = self._train_lists(k)

        draw = self._draw

        def fwd_bwd():
            if draw:
                self.rng.run(current_stream_ptr())
            fwd_list.run(current_stream_ptr())
            loss_call.run(current_stream_ptr(), join=False)   # side stream 1, beside the backward list ...
            bwd_list.run(current_stream_ptr())
            main = torch.cuda.current_stream()                # ... and joined explicitly: not every plan's backward
            ev = torch.cuda.Event()                           # list uses (hence joins) side stream 1
            ev.record(_side_stream(None, 1))
            main.wait_event(ev)

        def adam():
            optimizer.step_flat(m, gflat)

        if not m.use_cuda_graphs:
            fwd_bwd()
            if allreduce is not None:
                allreduce(gflat)
            adam()
        else:
            # the captured Adam launch bakes its hyper-parameters in: a changed lr / betas / eps / grad_scale
            # (LR scheduler, load_state_dict, a DataParallel wrapper created later) re-captures
            # FHVAE_DP_ONE_GRAPH=1: the collective is captured too (NCCL launches are capturable): one graph per step
            # instead of graph / eager all-reduce / graph
            one_graph = allreduce is not None and os.environ.get("FHVAE_DP_ONE_GRAPH", "0") == "1"
            key = ("train", alpha, id(optimizer), allreduce is None, optimizer.hyper_key(), draw, one_graph, v)
            graphs = self.__dict__.setdefault("_train_graphs", {})
            if key not in graphs:
                optimizer._state_for(m)               # allocate Adam state outside capture
                if allreduce is None:
                    graphs[key] = (self._capture(lambda: (fwd_bwd(), adam()), restore=optimizer), None)
                elif one_graph:
                    graphs[key] = (self._capture(lambda: (fwd_bwd(), allreduce(gflat), adam()), restore=optimizer), None)
                else:
                    graphs[key] = (self._capture(fwd_bwd), self._capture(adam, restore=optimizer))
            g1, g2 = graphs[key]
            g1.replay()
            if g2 is not None:
                allreduce(gflat)
                g2.replay()
        return self.loss

    # ------------------------------------------------------------------ data parallel, overlapped all-reduce
    def run_train_step_overlapped(self, optimizer, alpha, dp):
        """Data-parallel step whose gradient all-reduce overlaps the tail of the backward.  The flat gradient buffer is
        laid out [z2 encoder | z1 encoder + decoder + table]; the second range is complete when the z1 encoder's
        weight gradients are, i.e. before the z2 encoder's BPTT (the last ~15 % of the step) has run.  Three graphs:
        A = forward + backward down to the z1 BPTT launch (critical path), W = the z1 stack's weight-gradient / bias
        launches (side stream; followed there by the all-reduce of range 2), B = z2 head + z2 BPTT + its weight
        gradients (main stream, concurrent with W and the all-reduce).  Then the all-reduce of range 1 and Adam."""
        m = self.m
        k = 0
        gflat = m._grad_buffer(k)
        if self.bwd[k] is None or self.__dict__.get("_wgrad_split", True):
            # this schedule needs every z1 / decoder gradient complete before the z2 part starts: no deferred half
            self._wgrad_split = False
            self.bwd[k] = self._build_bwd(gflat)
            self.__dict__.pop("_train_list_cache", None)
            self.__dict__.pop("_train_graphs", None)
        if not hasattr(self, "loss"):
            self.loss = torch.zeros((), dtype=torch.float32, device=self.dev)
        B = self.B
        if self.__dict__.get("_gout_train") != (alpha, B):
            self.gout.zero_()
            self.gout[0].fill_(-1.0 / B)
            self.gout[5].fill_(-alpha / B)
            self._gout_train = (alpha, B)
            self._gout_rows = None
        draw = self._draw
        ov = self.__dict__.get("_ovl", {}).get(draw)
        if ov is None or ov["key"] != (alpha, id(optimizer), optimizer.hyper_key()):
            fwd, bwd = self._train_lists(k)
            jD = max(i for i, c_ in enumerate(bwd.calls) if c_[1] == "join" and c_[2] == 2)
            head = bwd.calls[:jD]
            iM = max(i for i, c_ in enumerate(head) if c_[0] is not None and c_[3] == 0)     # last main-stream launch
            tail = head[iM + 1:]
            assert all(c_[0] is None or c_[3] != 0 for c_ in tail)

            def mk(calls, keep):
                cl = CallList()
                cl.calls, cl.keep = list(calls), keep
                return cl
            lc = CallList()
            lc.add("fhvae_loss_mean", ptr(self.out), ptr(self.out, 5 * B), float(alpha), B, ptr(self.loss))
            A = mk((self.rng.calls if draw else []) + fwd.calls + lc.calls + head[:iM + 1] + [(None, "join", 2, 0)],
                   fwd.keep + bwd.keep)
            Wl = mk(tail, bwd.keep)
            Bl = mk(bwd.calls[jD + 1:], bwd.keep)
            optimizer._state_for(m)
            # one eager pass (module loading must not happen inside a capture; no segment is warmed up on its own, which
            # would double-apply accumulating kernels), then every piece is captured without a warm-up run
            for cl in (A, Wl, Bl):
                cl.run(current_stream_ptr())

            def cap(f):
                s_ = torch.cuda.Stream(priority=-1)
                s_.wait_stream(torch.cuda.current_stream())
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_, stream=s_):
                    f()
                return g_
            ov = self.__dict__.setdefault("_ovl", {})[draw] = dict(key=(alpha, id(optimizer), optimizer.hyper_key()),
                                  gA=cap(lambda: A.run(current_stream_ptr())), gW=cap(lambda: Wl.run(current_stream_ptr())),
                                  gB=cap(lambda: Bl.run(current_stream_ptr())),
                                  gC=cap(lambda: optimizer.step_flat(m, gflat)), side=torch.cuda.Stream())
        main, side = torch.cuda.current_stream(), ov["side"]
        z2_end = m._z2_end
        ov["gA"].replay()
        eA = torch.cuda.Event(); eA.record(main)
        with torch.cuda.stream(side):
            side.wait_event(eA)
            ov["gW"].replay()
            dp.allreduce_(gflat[z2_end:])                    # z1 encoder + decoder + table, beside the z2 BPTT
            eR = torch.cuda.Event(); eR.record(side)
        ov["gB"].replay()
        dp.allreduce_(gflat[:z2_end])
        main.wait_event(eR)
        ov["gC"].replay()
        return self.loss

    # ------------------------------------------------------------------ sharded mu2 table
    _DISC_FWD = ("fhvae_mu2_gather", "fhvae_disc_fwd_partial", "fhvae_disc_target", "fhvae_disc_combine")
    _DISC_BWD = ("fhvae_disc_bwd_segs", "fhvae_disc_bwd_rows", "fhvae_disc_bwd_finish", "fhvae_mu2_scatter_reduce")

    def _sharded_segments(self, k):
        """The train-step call lists cut where the discriminative chain meets the collectives:
        S1 forward up to the z2 posterior | S2 rest of the forward | S3 ELBO forward+backward seam | S4 backward
        down to the z2 head | S5 z2 head + z2 encoder backward.  The replicated-table kernels of the discriminative
        chain are dropped (run_train_step_sharded drives their sharded counterparts between the segments)."""
        cache = self.__dict__.setdefault("_shard_seg_cache", {})
        if k in cache:
            return cache[k]
        fwd, bwd = self._train_lists(k)
        nf = [c_[1] for c_ in fwd.calls]
        nb = [c_[1] for c_ in bwd.calls]
        iA = nf.index("fhvae_mu2_gather")
        assert tuple(nf[iA:iA + 4]) == self._DISC_FWD and nf[-1] in ("fhvae_elbo_fwd_bwd", "fhvae_elbo_fwd")
        fused = nf[-1] == "fhvae_elbo_fwd_bwd"
        jD = max(i for i, c_ in enumerate(bwd.calls) if c_[1] == "join" and c_[2] == 2)
        head = 0 if fused else 2                       # unfused: [step_coef, elbo_bwd] open the backward list
        assert sum(n_ in self._DISC_BWD for n_ in nb[:jD]) == 4, nb[:12]

        def seg(calls, keep):
            cl = CallList()
            cl.calls, cl.keep = list(calls), keep
            return cl
        s1 = seg(fwd.calls[:iA], fwd.keep)
        s2 = seg(fwd.calls[iA + 4:-1], fwd.keep)
        s3 = seg([fwd.calls[-1]] + (bwd.calls[:2] if not fused else []), fwd.keep + bwd.keep)
        s4 = seg([c_ for c_ in bwd.calls[head:jD] if c_[1] not in self._DISC_BWD], bwd.keep)
        s5 = seg(bwd.calls[jD + 1:], bwd.keep)
        cache[k] = (s1, s2, s3, s4, s5)
        return cache[k]

    def run_train_step_sharded(self, optimizer, alpha, dp):
        """One train step with the mu2 table sharded by row id over the ranks of ``dp`` (parallel.DataParallel):
        the model's ``mu2_table`` parameter holds rows ``rank, rank+W, ...``; ``self.idx`` are GLOBAL row ids.
        Exchange per step: all-gather (z2_mu, idx, g) | owner-served mu2 rows by reduce-scatter |
        all-gather of the (max, sumexp) partials -> rank-ordered LSE | all-to-all of the sum_n p_bn m_n partials |
        all-gather of the sparse row gradients, scatter-reduced by the owner in ascending global segment order |
        all-reduce of the dense gradient prefix.  The dense softmax part of d table stays owner-local."""
        m = self.m
        k = 0
        gflat = m._grad_buffer(k)
        if self.bwd[k] is None:
            self.bwd[k] = self._build_bwd(gflat)
        B, Z, W, rank = self.B, self.Z2, dp.world, dp.rank
        Bg = B * W
        if not hasattr(self, "loss"):
            self.loss = torch.zeros((), dtype=torch.float32, device=self.dev)
        if self.__dict__.get("_gout_train") != (alpha, B):
            self.gout.zero_()
            self.gout[0].fill_(-1.0 / B)
            self.gout[5].fill_(-alpha / B)
            self._gout_train = (alpha, B)
            self._gout_rows = None
        sh = self.__dict__.get("_shard")
        n_alloc = m.mu2_table.shape[0]
        if sh is None or sh["key"] != (W, rank, dp.n_local):
            f = self.f
            ns = _lib.fn("fhvae_disc_nsplit")(Bg, n_alloc)
            i64 = lambda n: torch.zeros(n, dtype=torch.int64, device=self.dev)
            sh = self._shard = dict(
                key=(W, rank, dp.n_local), ns=ns, packet=f(B, Z + 4), packet_g=f(Bg, Z + 4), lidx_g=i64(Bg), g_g=f(Bg),
                mu2_send=f(Bg, Z), part=f(ns, Bg, 2), part_all=f(W * ns, Bg, 2), lse_g=f(Bg), sumpm=f(ns, Bg, Z),
                sumpm_send=f(W, ns, B, Z), sumpm_recv=f(W * ns, B, Z), dmu2_g=f(Bg, Z),
                touched_g=torch.zeros(Bg, dtype=torch.int32, device=self.dev),
                stream=torch.cuda.Stream(), graphs={})
            self.touched_global = sh["touched_g"]
        ns = sh["ns"]
        segs = self._sharded_segments(k)
        lc = CallList()
        lc.add("fhvae_loss_mean", ptr(self.out), ptr(self.out, 5 * B), float(alpha), B, ptr(self.loss))
        if not hasattr(self, "coef"):
            raise RuntimeError("backward list not built")
        N_local = dp.n_local
        tab = ptr(m.mu2_table)
        dtab = ptr(gflat, m._off["mu2_table"])
        pk, pkg, ld = ptr(sh["packet"]), ptr(sh["packet_g"]), Z + 4
        z2h, ldz = ptr(self.z2head), 2 * Z
        g_loc = ptr(self.gout, 5 * B)
        call = lambda name, *a: _lib.check(_lib.fn(name)(*a, current_stream_ptr()), name)

        def run_seg(i, extra=None):
            if not segs[i].calls and extra is None:
                return

            def body():
                if i == 0 and draw:
                    self.rng.run(current_stream_ptr())
                segs[i].run(current_stream_ptr())
                if extra is not None:
                    extra.run(current_stream_ptr())
            # The first step runs eagerly (module loading / function attributes must not happen inside a capture, and a
            # warm-up replay of ONE segment would double-apply its accumulating kernels); from the second step on
            # each segment is captured once, without a warm-up run, and replayed.
            if not m.use_cuda_graphs or not sh.get("warm"):
                return body()
            key = ("seg", i, alpha, draw if i == 0 else None)
            if key not in sh["graphs"]:
                s_ = torch.cuda.Stream(priority=-1)
                s_.wait_stream(torch.cuda.current_stream())
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_, stream=s_):
                    body()
                sh["graphs"][key] = g_
            sh["graphs"][key].replay()

        draw = self._draw
        main, D = torch.cuda.current_stream(), sh["stream"]
        ev = lambda s: (lambda e: (e.record(s), e)[1])(torch.cuda.Event())

        run_seg(0)                                                   # ... z2 posterior
        eA = ev(main)
        with torch.cuda.stream(D):
            D.wait_event(eA)
            call("fhvae_shard_pack", z2h, ldz, ptr(self.idx), g_loc, pk, B, Z)
            dp.all_gather(sh["packet_g"], sh["packet"])
            call("fhvae_shard_unpack", pkg, Bg, Z, W, rank, N_local, None, ptr(sh["lidx_g"]), ptr(sh["g_g"]),
                 ptr(self.nan_flag))
            # owners serve the mu2 rows of ALL global segments; exactly one rank contributes a non-zero row
            sh["mu2_send"].zero_()
            call("fhvae_rows_copy", tab, ptr(sh["lidx_g"]), ptr(sh["mu2_send"]), None, Bg, Z)
            dp.reduce_scatter(self.mu2, sh["mu2_send"])
            if N_local > 0:
                call("fhvae_disc_fwd_partial", pkg, ld, tab, N_local, Z, ptr(sh["part"]), ns, Bg)
            else:
                sh["part"][..., 0].fill_(float("-inf")); sh["part"][..., 1].zero_()
            dp.all_gather(sh["part_all"], sh["part"])
            call("fhvae_disc_target", z2h, ldz, ptr(self.mu2), ptr(self.tgt), B, Z)
            call("fhvae_disc_combine_sharded", ptr(sh["part_all"]), W * ns, Bg, ptr(self.tgt), rank * B, B,
                 ptr(self.out, 5 * B), ptr(sh["lse_g"]))
            eB = ev(D)
            # backward pieces that only need (z2_mu, lse, g) of all segments: dense d table (owner-local), sum_n p_bn m_n
            if N_local > 0:
                call("fhvae_disc_bwd_rows", pkg, ld, tab, N_local, Z, ptr(sh["lse_g"]), ptr(sh["g_g"]), dtab, Bg)
                call("fhvae_disc_bwd_segs", pkg, ld, tab, N_local, Z, ptr(sh["lse_g"]), ptr(sh["sumpm"]), ns, Bg)
            else:
                sh["sumpm"].zero_()
            sh["sumpm_send"].copy_(sh["sumpm"].view(ns, W, B, Z).permute(1, 0, 2, 3))
            dp.all_to_all(sh["sumpm_recv"].view(W, ns, B, Z), sh["sumpm_send"])
        run_seg(1)                                                   # z1 encoder, decoder (beside the exchange)
        main.wait_event(eB)                                          # mu2 rows + log q(i|z2)
        run_seg(2, lc)                                               # ELBO forward/backward seam, loss
        eC = ev(main)
        with torch.cuda.stream(D):
            D.wait_event(eC)                                         # dz2head / dmu2 hold the ELBO part
            call("fhvae_disc_bwd_finish", z2h, ldz, ptr(self.mu2), ptr(sh["sumpm_recv"]), W * ns, g_loc,
                 ptr(self.dz2head), ldz, ptr(self.dmu2), B, Z)
            dp.all_gather(sh["dmu2_g"], self.dmu2)
            # sparse rows (KL + prior + target parts): the owner adds them in ascending GLOBAL segment order
            call("fhvae_mu2_scatter_reduce", ptr(sh["dmu2_g"]), ptr(sh["lidx_g"]), dtab, ptr(sh["touched_g"]), Bg, Z,
                 max(N_local, 1))
            eD = ev(D)
        run_seg(3)                                                   # decoder + z1 encoder backward
        main.wait_event(eD)
        run_seg(4)                                                   # z2 head + z2 encoder backward
        dense = gflat[:m._off["mu2_table"]]
        dp.allreduce_(dense)                                         # the table gradient stays with its owner

        def adam():
            optimizer.step_flat(m, gflat)
        if not m.use_cuda_graphs:
            adam()
        else:
            key = ("adam", id(optimizer), optimizer.hyper_key())
            if key not in sh["graphs"]:
                optimizer._state_for(m)
                sh["graphs"][key] = self._capture(adam, restore=optimizer)
            sh["graphs"][key].replay()
        sh["warm"] = True
        return self.loss

    def _train_lists(self, k):
        """Call lists of the fused train step.  The upstream gradients are known before the forward there, so the
        chain [elbo_fwd] ... [step_coef, elbo_bwd] at the forward/backward seam collapses into one
        ``fhvae_elbo_fwd_bwd`` launch (bit-identical results); everything else is shared with the autograd path."""
        cache = self.__dict__.setdefault("_train_list_cache", {})
        if k in cache:
            return cache[k]
        fwd, bwd = self.fwd, self.bwd[k]
        lists = (fwd, bwd)
        names = lambda cl, sl: [c_[1] for c_ in cl.calls[sl]]
        if (os.environ.get("FHVAE_FUSED_ELBO", "1") != "0" and names(fwd, slice(-1, None)) == ["fhvae_elbo_fwd"]
                and names(bwd, slice(0, 2)) == ["fhvae_step_coef", "fhvae_elbo_bwd"]):
            fa, ca, ba = fwd.calls[-1][2], bwd.calls[0][2], bwd.calls[1][2]
            # elbo_fwd: (x, xhead, xs_b, xs_t, lv_off, z1head, z2head, mu2, nsegs, out, nan_flag, B, T, F, Z1, Z2)
            # step_coef: (gout, nsegs, coef, detach_px, prior_grad, B)
            # elbo_bwd: (x, xhead, xs_b, xs_t, lv_off, z1head, z2head, mu2, coef, dxhead, dz1head, dz2head, dmu2, B, ...)
            fused = (_lib.fn("fhvae_elbo_fwd_bwd"), "fhvae_elbo_fwd_bwd",
                     fa[:9] + (ca[0], ca[3], ca[4]) + fa[9:11] + ba[9:13] + fa[11:], 0)
            ft, bt = CallList(), CallList()
            ft.calls, ft.keep = fwd.calls[:-1] + [fused], fwd.keep
            bt.calls, bt.keep = bwd.calls[2:], bwd.keep
            lists = (ft, bt)
        cache[k] = lists
        return lists

    def _capture(self, f, restore=None):
        # the warm-up run below really executes f once: snapshot what an optimizer step would change
        snap = None
        if restore is not None:
            st = restore._state_for(self.m)
            snap = (self.m._flat.clone(), st["m"].clone(), st["v"].clone(), st["step"].clone())
        g = self._capture_inner(f)
        if snap is not None:
            st = restore._state_for(self.m)
            self.m._flat.copy_(snap[0]); st["m"].copy_(snap[1]); st["v"].copy_(snap[2]); st["step"].copy_(snap[3])
        return g

    def _capture_inner(self, f):
        # warm up on a side stream (lazy module loading must not happen inside capture), then capture.
        # The capture stream has HIGH priority: the critical-path kernels (cluster recurrence) then win
        # the SMs over the weight-gradient GEMMs that CallList forks onto its default-priority stream.
        s = torch.cuda.Stream(priority=-1)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            f()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            f()
        return g

    # ---- shared tail: gather + discriminative term (side stream 2: needs only z2head) + ELBO
    def _disc_fwd(self):
        c, m, B = self.fwd, self.m, self.B
        tab = ptr(m.mu2_table)
        c.add("fhvae_mu2_gather", tab, ptr(self.idx), ptr(self.mu2), B, self.Z2, self.N, ptr(self.nan_flag), side=2)
        c.add("fhvae_disc_fwd_partial", ptr(self.z2head), 2 * self.Z2, tab, self.N, self.Z2,
              ptr(self.part), self.nsplit, B, side=2)
        c.add("fhvae_disc_target", ptr(self.z2head), 2 * self.Z2, ptr(self.mu2), ptr(self.tgt), B, self.Z2, side=2)
        c.add("fhvae_disc_combine", ptr(self.part), self.nsplit, ptr(self.tgt), ptr(self.out, 5 * B),
              ptr(self.lse), B, side=2)

    def _tail_fwd(self, xhead, xs_b, xs_t, lv_off, disc=True):
        c, B = self.fwd, self.B
        if disc:
            self._disc_fwd()
        c.join(2)                                    # mu2 rows (gather) for the KL(z2 || mu2) term
        c.add("fhvae_elbo_fwd", ptr(self.x), ptr(xhead), xs_b, xs_t, lv_off, ptr(self.z1head),
              ptr(self.z2head), ptr(self.mu2), ptr(self.nsegs), ptr(self.out), ptr(self.nan_flag),
              B, self.T, self.F, self.Z1, self.Z2)

    def _tail_bwd(self, c: CallList, gflat, xhead, dxhead, xs_b, xs_t, lv_off, disc=True):
        """coef from upstream grads; ELBO bwd; disc bwd; table gradient (dense + sparse rows).  ``disc=False``: the
        caller issues the discriminative chain itself (``_disc_bwd``) at a later point of the list."""
        m, B, Z2 = self.m, self.B, self.Z2
        f = self.f
        if not hasattr(self, "coef"):
            self.coef = f(4, B)
            self.dz1head, self.dz2head = f(B, 2 * self.Z1), f(B, 2 * Z2)
            self.dmu2 = f(B, Z2)
            self.sumpm = f(self.nsplit, B, Z2)
            self.touched = torch.zeros(B, dtype=torch.int32, device=self.dev)
            self.nsegs_f = f(B)
        gout, coef = self.gout, self.coef
        detach_px, prior_grad = m.detach_px, m.prior_grad

        c.add("fhvae_step_coef", ptr(gout), ptr(self.nsegs), ptr(coef), int(detach_px), int(prior_grad), B)
        dtab = ptr(gflat, m._off["mu2_table"])
        tab = ptr(m.mu2_table)
        c.add("fhvae_elbo_bwd", ptr(self.x), ptr(xhead), xs_b, xs_t, lv_off, ptr(self.z1head),
              ptr(self.z2head), ptr(self.mu2), ptr(coef), ptr(dxhead), ptr(self.dz1head),
              ptr(self.dz2head), ptr(self.dmu2), B, self.T, self.F, self.Z1, Z2)
        if disc:
            self._disc_bwd(c, gflat)

    def _disc_bwd(self, c: CallList, gflat):
        m, B, Z2 = self.m, self.B, self.Z2
        dtab = ptr(gflat, m._off["mu2_table"])
        tab = ptr(m.mu2_table)
        g_qy = ptr(self.gout, 5 * B)
        # the discriminative / table-gradient chain only meets the main sequence again at the z2 head: side stream 2
        c.add("fhvae_disc_bwd_segs", ptr(self.z2head), 2 * Z2, tab, self.N, Z2, ptr(self.lse),
              ptr(self.sumpm), self.nsplit, B, side=2)
        c.add("fhvae_disc_bwd_rows", ptr(self.z2head), 2 * Z2, tab, self.N, Z2, ptr(self.lse), g_qy,
              dtab, B, side=2)
        c.add("fhvae_disc_bwd_finish", ptr(self.z2head), 2 * Z2, ptr(self.mu2), ptr(self.sumpm),
              self.nsplit, g_qy, ptr(self.dz2head), 2 * Z2, ptr(self.dmu2), B, Z2, side=2)
        c.add("fhvae_mu2_scatter_reduce", ptr(self.dmu2), ptr(self.idx), dtab, ptr(self.touched), B, Z2,
              self.N, side=2)

    def _head_fwd(self, c, srcs, wname, bname, head, Z):
        """head (B,2Z) = sum_l src_l @ W[:, cols_l]^T + b, W = [mulayer.weight ; logvar_layer.weight]."""
        m, B = self.m, self.B
        ldw = sum(k for _, _, k in srcs)
        col = 0
        for i, (src, lds, K) in enumerate(srcs):
            c.gemm([gemm_nt(src, lds, m.poff(wname, col), ldw, ptr(head), 2 * Z, B, 2 * Z, K,
                            bias=m.poff(bname) if i == 0 else 0, beta=0.0 if i == 0 else 1.0)], self.mode)
            col += K


def _lstm_names(prefix: str, l: int):
    return (f"{prefix}.lstm.weight_ih_l{l}", f"{prefix}.lstm.weight_hh_l{l}",
            f"{prefix}.lstm.bias_ih_l{l}", f"{prefix}.lstm.bias_hh_l{l}")


class _FHVAEPlan(_Plan):
    NETS = (("z2", "z2_pre_encoder"), ("z1", "z1_pre_encoder"), ("dec", "pre_decoder"))

    def __init__(self, m: "FHVAE", B, T, F):
        super().__init__(m, B, T, F)
        f, Z1, Z2 = self.f, self.Z1, self.Z2
        self.H = {"z2": m.z2_hus[0], "z1": m.z1_hus[0], "dec": m.x_hus[0]}
        self.L = {"z2": len(m.z2_hus), "z1": len(m.z1_hus), "dec": len(m.x_hus)}
        self.x_tm = f(T, B, F)
        self.bsum = f(m._bias_block_len)
        self.h, self.c, self.acts, self.P = {}, {}, {}, {}
        # two-layer stacks run as ONE layer-wavefront launch (lstm_wave.cu): no layer-1 projection GEMM / buffer
        use_wave = os.environ.get("FHVAE_WAVE", "1") != "0"
        self.wave = {k: bool(use_wave and self.L[k] == 2 and len(set(hus)) == 1 and
                             _lib.fn("fhvae_lstm_wave_supported")(T, B, self.H[k], 2, self.mode))
                     for (k, _), hus in zip(self.NETS, (m.z2_hus, m.z1_hus, m.x_hus))}
        self.wave_xchg, self.wave_xchg_bwd = {}, {}      # one exchange buffer per hidden width (zeroed ONCE)
        # pre-packed weight operands of the wavefront kernels (fhvae_lstm_wave_pack): rebuilt at the start of every
        # forward list from the current weights, consumed by the stack's forward AND BPTT launch
        self.wave_packed = {}
        if os.environ.get("FHVAE_WAVE_PACK", "1") != "0":
            for k, on in self.wave.items():
                if on:
                    nb = _lib.fn("fhvae_lstm_wave_pack_bytes")(self.H[k], 2, self.mode)
                    self.wave_packed[k] = torch.zeros(nb // 4, dtype=torch.float32, device=self.dev)
        for k, on in self.wave.items():
            if on and self.H[k] not in self.wave_xchg:
                nbytes = _lib.fn("fhvae_lstm_wave_xchg_bytes")(T, B, self.H[k], 2)
                self.wave_xchg[self.H[k]] = torch.zeros(nbytes // 4, dtype=torch.float32, device=self.dev)
        for k, _ in self.NETS:
            H = self.H[k]
            for l in range(self.L[k]):
                self.h[k, l], self.c[k, l] = f(T, B, H), f(T, B, H)
                self.acts[k, l] = f(T, B, 4 * H)
                if not (k == "dec" and l == 0) and not (self.wave[k] and l == 1):
                    self.P[k, l] = f(T, B, 4 * H)
        # Long-K weight gradients (K = T*B rows) run on the TMA-fed kernel from bf16 hi/lo planes (gemm_wgrad.cu).
        # Wavefront stacks write the planes of h (forward) and dgates (backward) themselves; everything else
        # (x, non-wave stacks, the decoder-head gradient) goes through fhvae_split_planes_batch.
        self.tma_wgrad = (self.mode != _lib.MODE_F32_SIMT and os.environ.get("FHVAE_TMA_WGRAD", "1") != "0"
                          and F % 8 == 0 and all(h % 8 == 0 for h in self.H.values()) and (T - 1) * B >= 1024)
        self.planes = {}
        # Optional (FHVAE_TMA_PROJ=1): the short-K products on the critical path (layer-0 projections of x, the decoder head)
        # on the TMA-fed K-major kernel (gemm_proj.cu): operands = bf16 hi/lo planes of x_tm / h_dec and of the weights (one
        # split launch at the start of the step).  Measured (tools/proj_bench.py, warm): x projection 21.5 us vs 16.6 us for
        # gemm_tc, decoder head 12.2 vs 14.4 us; in the step 0.918 vs 0.906 ms -- with ~1 tile per CTA both kernels are bound
        # by the per-CTA latency chain, not by the operand path, so the default stays gemm_tc.
        self.tma_proj = (self.mode != _lib.MODE_F32_SIMT and os.environ.get("FHVAE_TMA_PROJ", "0") == "1" and F % 16 == 0
                         and all(h % 16 == 0 for h in self.H.values()))
        self.wplanes = {}
        self.Q = {"z1": f(B, 4 * self.H["z1"]), "dec": f(B, 4 * self.H["dec"])}
        self.xchg = f(16, B, max(self.H.values()))      # L2-resident exchange scratch of the cluster kernels
        self.xhead = f(T, B, 2 * F)
        self._build_fwd()

    def planes_of(self, t: torch.Tensor) -> torch.Tensor:
        """bf16 (2, numel) hi/lo planes that shadow the fp32 tensor t (allocated on first use)."""
        key = t.data_ptr()
        if key not in self.planes:
            self.planes[key] = (torch.zeros(2, t.numel(), dtype=torch.bfloat16, device=self.dev), t)
        return self.planes[key][0]

    def px_views(self):
        F = self.F
        xh = self.xhead.permute(1, 0, 2)[:self.valid]    # (B,T,2F) view of the time-major buffer
        return [xh[..., :F], xh[..., F:]]

    def _bs(self, k, l):                             # fused (b_ih + b_hh) pointer
        return ptr(self.bsum, self.m._bias_off[k, l])

    def _build_fwd(self):
        c, m, B, T, F = self.fwd, self.m, self.B, self.T, self.F
        Z1, Z2, mode = self.Z1, self.Z2, self.mode
        TB = T * B
        pre = dict(self.NETS)
        for bi, bh, n, boff in m._bias_blocks:       # fused (b_ih + b_hh), beside the transpose
            c.add("fhvae_add2", ptr(self.bsum, boff), m.poff(bi), m.poff(bh), n, side=2)
        # weight operand images: one side stream per stack (3, 5, 6), so that all three are built concurrently at the very
        # start of the step and the first wavefront launch never waits for an image it does not need
        self._pack_side = {}
        for i, (k, buf) in enumerate(self.wave_packed.items()):
            _, whh0, _, _ = _lstm_names(dict(self.NETS)[k], 0)
            wih1, whh1, _, _ = _lstm_names(dict(self.NETS)[k], 1)
            self._pack_side[k] = (3, 5, 6)[i] if os.environ.get("FHVAE_PACK_STREAMS", "3") == "3" else 3
            c.add("fhvae_lstm_wave_pack", m.poff(whh0), m.poff(wih1), m.poff(whh1), ptr(buf), self.H[k], 2, mode,
                  side=self._pack_side[k])
        Hz2, Hz1, Hd = self.H["z2"], self.H["z1"], self.H["dec"]
        wih_z2, _, _, _ = _lstm_names(pre["z2"], 0)
        wih_z1, _, _, _ = _lstm_names(pre["z1"], 0)
        c.join(2)          # (x_tm, the time-major copy of x, is written together with x by _Plan.set_x)
        # layer-0 projections of x: the z2 encoder's is on the critical path, the z1 encoder's overlaps the
        # z2 recurrence on side stream 1 (joined before the z1 stack).  (Running them on the TMA-fed kernel over
        # feature-major planes of x was tried: 19.4 us vs 22 us -- both are bound by the 21 MB output, not worth two
        # more operand-preparation kernels.)
        Hd_, Ld_ = self.H["dec"], self.L["dec"]
        self.proj_head = bool(self.tma_proj and self.wave["dec"] and self.tma_wgrad)   # h_dec planes come from the wave kernel
        if self.tma_proj:
            bf = lambda rows, k: torch.zeros(2, rows * k, dtype=torch.bfloat16, device=self.dev)
            self.wplanes = {"z2": bf(4 * Hz2, F), "z1": bf(4 * Hz1, F), "dec": bf(2 * F, Hd_)}
            xp = self.planes_of(self.x_tm)
            sp = [SplitProblem(ptr(self.x_tm), xp.data_ptr(), F, F, TB * F, TB, F),
                  SplitProblem(m.poff(wih_z2), self.wplanes["z2"].data_ptr(), F, F, 4 * Hz2 * F, 4 * Hz2, F),
                  SplitProblem(m.poff(wih_z1), self.wplanes["z1"].data_ptr(), F + Z2, F, 4 * Hz1 * F, 4 * Hz1, F),
                  SplitProblem(m.poff("dec_gauss_layer.mulayer.weight"), self.wplanes["dec"].data_ptr(), Hd_, Hd_,
                               2 * F * Hd_, 2 * F, Hd_)]
            arr = (SplitProblem * len(sp))(*sp)
            c.keep.append(arr)
            c.add("fhvae_split_planes_batch", arr, len(sp))
            pj = lambda k, H_: ProjProblem(xp.data_ptr(), self.wplanes[k].data_ptr(), ptr(self.P[k, 0]), self._bs(k, 0),
                                           TB, 4 * H_, F, 0, F, TB * F, F, 4 * H_ * F, 4 * H_)
            c.proj([pj("z2", Hz2)], mode)
            c.proj([pj("z1", Hz1)], mode, side=1)
        else:
            c.gemm([gemm_nt(ptr(self.x_tm), F, m.poff(wih_z2), F, ptr(self.P["z2", 0]), 4 * Hz2, TB, 4 * Hz2, F,
                            bias=self._bs("z2", 0))], mode)
            # The z1 encoder's projection runs beside the z2 recurrence (side stream 1).  It is forked right BEFORE the z2
            # wavefront launch, behind the same dependencies: both become ready together, the wavefront (high-priority
            # stream) takes its 128 SMs first and the 320 GEMM CTAs trickle in on the rest.  (Forked right behind the z2
            # projection it started first and held the SMs: the wavefront launch became resident ~10 us late.)
            z1_proj = lambda: c.gemm([gemm_nt(ptr(self.x_tm), F, m.poff(wih_z1), F + Z2, ptr(self.P["z1", 0]), 4 * Hz1, TB,
                                              4 * Hz1, F, bias=self._bs("z1", 0))], mode, side=1)
            if not self.wave["z2"] or os.environ.get("FHVAE_Z1PROJ_LATE", "1") == "0":
                z1_proj()
                z1_proj = None

        def stack(k, q0):
            H = self.H[k]
            if self.wave[k]:
                _, whh0, _, _ = _lstm_names(pre[k], 0)
                wih1, whh1, _, _ = _lstm_names(pre[k], 1)
                hp = [self.planes_of(self.h[k, l]).data_ptr() if self.tma_wgrad else None for l in (0, 1)]
                c.add("fhvae_lstm_wave_fwd_planes", ptr(self.P[k, 0]) if (k, 0) in self.P else None, q0, m.poff(whh0),
                      ptr(self.h[k, 0]), ptr(self.c[k, 0]), ptr(self.acts[k, 0]), m.poff(wih1), self._bs(k, 1),
                      m.poff(whh1), ptr(self.h[k, 1]), ptr(self.c[k, 1]), ptr(self.acts[k, 1]),
                      ptr(self.wave_xchg[H]), hp[0], hp[1], T * B * H, ptr(self.wave_packed[k]) if k in self.wave_packed else None,
                      T, B, H, 2, mode)
                return
            for l in range(self.L[k]):
                wih, whh, _, _ = _lstm_names(pre[k], l)
                if l > 0:
                    c.gemm([gemm_nt(ptr(self.h[k, l - 1]), H, m.poff(wih), H, ptr(self.P[k, l]), 4 * H, TB,
                                    4 * H, H, bias=self._bs(k, l))], mode)
                Pp = ptr(self.P[k, l]) if (k, l) in self.P else None
                Qp = q0 if l == 0 else None
                c.add("fhvae_lstm_fwd", Pp, Qp, m.poff(whh), ptr(self.h[k, l]), ptr(self.c[k, l]),
                      ptr(self.acts[k, l]), ptr(self.xchg), T, B, H, mode)

        def final_h(k):
            H = self.H[k]
            return [(ptr(self.h[k, l], (T - 1) * B * H), H, H) for l in range(self.L[k])]

        wih_d, _, _, _ = _lstm_names(pre["dec"], 0)
        self.fused_heads = (max(self.L["z2"], self.L["z1"]) <= 2 and 2 * max(Z1, Z2) <= 128 and Z1 + Z2 <= 128
                            and max(self.L["z2"] * Hz2, self.L["z1"] * Hz1, 4 * Hz1, 4 * Hd) <= 1024
                            and os.environ.get("FHVAE_FUSED_HEADS", "1") != "0")

        def head_stage(k, H, wname, bname, head, Z, eps, zoff, Wq, ld_wq, bq, qoff, Kq, Q, NQ):
            """Gaussian head on the final hidden states -> sample -> hoisted projection for the next stack."""
            if self.fused_heads:
                src = [ptr(self.h[k, l], (T - 1) * B * H) for l in range(self.L[k])] + [None]
                c.add("fhvae_head_fwd", src[0], src[1], H, self.L[k], H, m.poff(wname), m.poff(bname), ptr(head), Z,
                      ptr(eps), ptr(self.zcat), Z1 + Z2, zoff, Wq, ld_wq, bq, qoff, Kq, ptr(Q), NQ, B)
                return
            self._head_fwd(c, final_h(k), wname, bname, head, Z)
            c.add("fhvae_reparam_fwd", ptr(head), 2 * Z, ptr(eps), ptr(self.zcat, zoff), Z1 + Z2, B, Z)
            c.gemm([gemm_nt(ptr(self.zcat, qoff), Z1 + Z2, Wq, ld_wq, ptr(Q), NQ, B, NQ, Kq, bias=bq or 0)], mode)

        # z2 encoder -> head -> sample (into zcat[:, Z1:]) -> time-invariant z2 part of the z1 encoder's input (Q)
        for k in ("z2",) if "z2" in self.wave_packed else ():
            c.join(self._pack_side[k])          # this stack's packed weight operands
        if not self.tma_proj and z1_proj is not None:
            z1_proj()
        stack("z2", None)
        head_stage("z2", Hz2, "z2_gauss_layer.mulayer.weight", "z2_gauss_layer.mulayer.bias", self.z2head, Z2,
                   self.eps2, Z1, m.poff(wih_z1, F), F + Z2, None, Z1, Z2, self.Q["z1"], 4 * Hz1)
        self._disc_fwd()       # log q(i|z2) only needs the z2 posterior: side stream 2, beside the z1 / decoder stacks
        c.join(1)
        if "z1" in self.wave_packed:
            c.join(self._pack_side["z1"])
        stack("z1", ptr(self.Q["z1"]))
        # z1 head -> sample -> the decoder's whole (time-invariant) layer-0 input projection
        head_stage("z1", Hz1, "z1_gauss_layer.mulayer.weight", "z1_gauss_layer.mulayer.bias", self.z1head, Z1,
                   self.eps1, 0, m.poff(wih_d), Z1 + Z2, self._bs("dec", 0), 0, Z1 + Z2, self.Q["dec"], 4 * Hd)
        self.n_encode_calls = len(c.calls)
        if "dec" in self.wave_packed:
            c.join(self._pack_side["dec"])
        stack("dec", ptr(self.Q["dec"]))
        Ld = self.L["dec"]
        if self.proj_head:
            hp = self.planes_of(self.h["dec", Ld - 1])
            c.proj([ProjProblem(hp.data_ptr(), self.wplanes["dec"].data_ptr(), ptr(self.xhead),
                                m.poff("dec_gauss_layer.mulayer.bias"), TB, 2 * F, Hd, 0, Hd, TB * Hd, Hd, 2 * F * Hd, 2 * F)], mode)
        else:
            c.gemm([gemm_nt(ptr(self.h["dec", Ld - 1]), Hd, m.poff("dec_gauss_layer.mulayer.weight"), Hd,
                            ptr(self.xhead), 2 * F, TB, 2 * F, Hd,
                            bias=m.poff("dec_gauss_layer.mulayer.bias"))], mode)
        self._tail_fwd(self.xhead, 2 * F, B * 2 * F, F, disc=False)

    def _build_bwd(self, gflat) -> CallList:
        c, m, B, T, F = CallList(), self.m, self.B, self.T, self.F
        Z1, Z2, mode, f = self.Z1, self.Z2, self.mode, self.f
        TB = T * B
        pre = dict(self.NETS)
        g = lambda name, extra=0: ptr(gflat, m._off[name] + extra)
        if not hasattr(self, "dxhead"):
            Hmax = max(self.H.values())
            self.dxhead = f(T, B, 2 * F)
            self.dhA, self.dhB = f(T, B, Hmax), f(T, B, Hmax)
            self.dhT = {(k, l): f(B, self.H[k]) for k in ("z1", "z2") for l in range(self.L[k])}
            self.dg, self.dgsum = {}, {}
            for k, _ in self.NETS:
                for l in range(self.L[k]):
                    self.dg[k, l] = f(T, B, 4 * self.H[k])
                    self.dgsum[k, l] = f(B, 4 * self.H[k])
            self.dzcat = f(B, Z1 + Z2)
            self.dh_rec, self.dc = self.xchg, f(B, Hmax)
        # (the discriminative chain is forked right before the decoder's BPTT launch, see below)
        late_disc = os.environ.get("FHVAE_DISC_LATE", "1") != "0" and not m.detach_px
        self._tail_bwd(c, gflat, self.xhead, self.dxhead, 2 * F, B * 2 * F, F, disc=not late_disc)
        cs: List = []          # bias column sums (one grouped side launch at the end)

        use_tma = self.tma_wgrad
        planes_of = self.planes_of
        made = set()           # fp32 tensors whose planes are written by the kernel that produces them

        def split(ts, side=1):   # (side: the stream of the launch that consumes the planes)
            probs = [SplitProblem(ptr(t), planes_of(t).data_ptr(), t.shape[-1], t.shape[-1], t.numel(),
                                  t.numel() // t.shape[-1], t.shape[-1]) for t in ts]
            for i in range(0, len(probs), _lib.SPLIT_MAX_BATCH):
                chunk = probs[i:i + _lib.SPLIT_MAX_BATCH]
                arr = (SplitProblem * len(chunk))(*chunk)
                c.keep.append(arr)
                c.add("fhvae_split_planes_batch", arr, len(chunk), side=side)

        wgp: List = []         # WgradProblem (TMA path)
        fresh: List = []       # tensors to split right before the next flush (gradients produced by the last launch)

        def wgrad_tn(G, g_row, X, x_row, Cp, ldc, M, N, K, fresh_g=True):
            """dW (M,N) = G[g_row : g_row+K]^T @ X[x_row : x_row+K]  (G (rows,M), X (rows,N) contiguous fp32)."""
            if use_tma and K >= 1024:
                pg, px = planes_of(G), planes_of(X)
                if fresh_g and G.data_ptr() not in made and all(G is not t for t in fresh):
                    fresh.append(G)
                wgp.append(WgradProblem(pg.data_ptr() + 2 * g_row * M, px.data_ptr() + 2 * x_row * N, Cp, M, N, K, 0,
                                        M, G.numel(), N, X.numel(), ldc))
            else:
                wg.append(gemm_tn(ptr(G, g_row * M), M, ptr(X, x_row * N), N, Cp, ldc, M, N, K))

        class _Side(list):     # weight-gradient GEMMs + bias column sums: flushed to side stream 1 right behind their inputs
            def flush(self_, side=1):
                if fresh:
                    split(list(fresh), side=side)
                    fresh.clear()
                for i in range(0, len(wgp), _lib.WGRAD_MAX_BATCH):
                    chunk = wgp[i:i + _lib.WGRAD_MAX_BATCH]
                    arr = (WgradProblem * len(chunk))(*chunk)
                    c.keep.append(arr)
                    c.add("fhvae_wgrad_planes_batch", arr, len(chunk), mode, side=side)
                wgp.clear()
                if self_:
                    c.gemm(list(self_), mode, side=side)
                    self_.clear()
                if cs:
                    # tiny, latency-bound: next to, not behind, the GEMMs -- and on its own stream, so that the join
                    # of the discriminative chain (side 2) at the z2 head never waits for a column sum
                    c.colsum(list(cs), side=3)
                    cs.clear()
        wg = _Side()
        if use_tma:    # (the planes of x_tm already exist when the forward projections ran on the TMA kernel)
            split([self.h[k, l] for k, _ in self.NETS for l in range(self.L[k]) if not self.wave[k]] +
                  ([] if self.tma_proj else [self.x_tm]))

        split_mode = int(os.environ.get("FHVAE_WGRAD_SPLIT", "2"))     # 0: one launch behind the BPTT; 1: layer-1 half there,
        split_wgrad = split_mode != 0 and self.__dict__.get("_wgrad_split", True)   # layer-0 half deferred; 2: all deferred

        def stack_bwd(k, dh_all_top, dh_last_of, extra=None, defer=None):
            """BPTT through the stack of net k, top layer first.  Returns nothing; fills dg/dgsum.  ``extra`` appends
            the weight gradients that only need this stack's dgates, so that they share its side-stream flush."""
            H = self.H[k]
            dh_all = dh_all_top
            if self.wave[k]:
                # both layers in one wavefront launch; dh of layer 0 never exists in HBM (lstm_wave.cu)
                if H not in self.wave_xchg_bwd:
                    nbytes = _lib.fn("fhvae_lstm_wave_bwd_xchg_bytes")(T, B, H, 2)
                    self.wave_xchg_bwd[H] = torch.zeros(nbytes // 4, dtype=torch.float32, device=self.dev)
                n0, n1 = _lstm_names(pre[k], 0), _lstm_names(pre[k], 1)
                dgp = [planes_of(self.dg[k, l]).data_ptr() if use_tma else None for l in (0, 1)]
                if use_tma:
                    made.update(self.dg[k, l].data_ptr() for l in (0, 1))
                # with the planes as output the fp32 dgates are never read again: not written at all
                c.add("fhvae_lstm_wave_bwd_planes", dh_all_top, dh_last_of(1), dh_last_of(0), m.poff(n1[1]),
                      ptr(self.c[k, 1]), ptr(self.acts[k, 1]), None if use_tma else ptr(self.dg[k, 1]),
                      ptr(self.dgsum[k, 1]), m.poff(n1[0]), m.poff(n0[1]), ptr(self.c[k, 0]), ptr(self.acts[k, 0]),
                      None if use_tma else ptr(self.dg[k, 0]), ptr(self.dgsum[k, 0]), ptr(self.wave_xchg_bwd[H]),
                      dgp[1], dgp[0], T * B * 4 * H, ptr(self.wave_packed[k]) if k in self.wave_packed else None,
                      T, B, H, 2, mode)
                # A full-width weight-gradient launch (128 CTAs x 193 KB) right behind the BPTT holds the SMs the NEXT
                # stack's 128-CTA wavefront launch needs until it has drained (CUPTI: the next BPTT became fully resident
                # ~28-33 us after its head kernel had finished).  So only the layer-1 half is launched here, where it hides
                # behind the head kernel; the layer-0 half is forked AFTER the head kernel (`defer`): it becomes ready
                # together with the next wavefront launch, which wins the SMs (high-priority stream), and trickles in on
                # the ~20 SMs the wavefront leaves free.  The last stack has nothing to hide behind: one launch.
                for l in (1, 0):
                    wih, whh, bih, bhh = _lstm_names(pre[k], l)
                    if T > 1:
                        wgrad_tn(self.dg[k, l], B, self.h[k, l], 0, g(whh), H, 4 * H, H, (T - 1) * B)
                    else:
                        c.torch_op(lambda o=m._off[whh], n=4 * H * H: gflat[o:o + n].zero_())
                    cs.append(ColsumProblem(ptr(self.dgsum[k, l]), g(bih), g(bhh), 4 * H, B, 4 * H))
                    if l > 0:
                        wgrad_tn(self.dg[k, l], 0, self.h[k, l - 1], 0, g(wih), H, 4 * H, H, TB)
                        if defer is not None and split_mode == 1:
                            wg.flush()                  # layer-1 half + every column sum queued so far
                if extra:
                    extra()
                if defer is None:
                    wg.flush()
                else:                                   # layer-0 half: flushed by the caller behind the next head kernel
                    held = (list(wgp), list(wg), list(fresh))
                    wgp.clear(); wg.clear(); fresh.clear()
                    if split_mode != 1 and cs:          # (column sums are never deferred)
                        c.colsum(list(cs), side=3)
                        cs.clear()

                    def later():
                        wgp.extend(held[0]); wg.extend(held[1]); fresh.extend(held[2])
                        wg.flush(side=4)
                    defer.append(later)
                return
            for l in reversed(range(self.L[k])):
                wih, whh, bih, bhh = _lstm_names(pre[k], l)
                c.add("fhvae_lstm_bwd", dh_all, dh_last_of(l), m.poff(whh), ptr(self.c[k, l]),
                      ptr(self.acts[k, l]), ptr(self.dg[k, l]), ptr(self.dgsum[k, l]), ptr(self.dh_rec),
                      ptr(self.dc), T, B, H, mode)
                # dW_hh = dg[1:]^T @ h[:-1]
                if T > 1:
                    wgrad_tn(self.dg[k, l], B, self.h[k, l], 0, g(whh), H, 4 * H, H, (T - 1) * B)
                else:
                    c.torch_op(lambda o=m._off[whh], n=4 * H * H: gflat[o:o + n].zero_())
                cs.append(ColsumProblem(ptr(self.dgsum[k, l]), g(bih), g(bhh), 4 * H, B, 4 * H))
                if l > 0:
                    wgrad_tn(self.dg[k, l], 0, self.h[k, l - 1], 0, g(wih), H, 4 * H, H, TB)
                    nxt = self.dhA if dh_all != ptr(self.dhA) else self.dhB
                    c.gemm([gemm_nn(ptr(self.dg[k, l]), 4 * H, m.poff(wih), H, ptr(nxt), H, TB, H, 4 * H)], mode)
                    dh_all = ptr(nxt)
                if extra and l == 0:
                    extra()
                wg.flush()

        def head_bwd(k, dhead, Z, wname, bname, eps=None, head=None, roff=0, dgsum=None, NG=0, Wq=None, ld_wq=0,
                     Kq=0, dzoff=0, beta=0):
            """Backward of a latent stage: d(sample) from the consumer stack's hoisted projection (dgsum @ Wq), the
            reparameterisation into dhead, dW/db of the Gaussian head and the dh_last its final hidden states get."""
            H, L = self.H[k], self.L[k]
            if self.fused_heads:
                dh = [ptr(self.dhT[k, l]) for l in range(L)] + [None]
                c.add("fhvae_head_bwd", dgsum, NG, Wq, ld_wq, Kq, ptr(self.dzcat), Z1 + Z2, dzoff,
                      beta, ptr(head) if eps is not None else None, ptr(eps) if eps is not None else None, Z, roff,
                      ptr(dhead), 1, m.poff(wname), L, H, dh[0], dh[1], B)
            else:
                if dgsum:
                    c.gemm([gemm_nn(dgsum, NG, Wq, ld_wq, ptr(self.dzcat, dzoff), Z1 + Z2, B, Kq, NG,
                                    beta=float(beta))], mode)
                if eps is not None:
                    c.add("fhvae_reparam_bwd", ptr(head), 2 * Z, ptr(eps), ptr(self.dzcat, roff), Z1 + Z2,
                          ptr(dhead), 2 * Z, 1, B, Z)
                c.gemm([gemm_nn(ptr(dhead), 2 * Z, m.poff(wname, l * H), L * H, ptr(self.dhT[k, l]), H, B, H, 2 * Z)
                        for l in range(L)], mode)
            cs.append(ColsumProblem(ptr(dhead), g(bname), None, 2 * Z, B, 2 * Z))
            for l in range(L):
                hT = ptr(self.h[k, l], (T - 1) * B * H)
                wg.append(gemm_tn(ptr(dhead), 2 * Z, hT, H, g(wname, l * H), L * H, 2 * Z, H, B))
            wg.flush()         # everything queued so far is ready: runs beside this stack's BPTT, not after it

        deferred: List = []
        # ---------------- decoder
        Hd, Ld = self.H["dec"], self.L["dec"]
        if not m.detach_px:
            wgrad_tn(self.dxhead, 0, self.h["dec", Ld - 1], 0, g("dec_gauss_layer.mulayer.weight"), Hd, 2 * F, Hd, TB)
            cs.append(ColsumProblem(ptr(self.dxhead), g("dec_gauss_layer.mulayer.bias"), None, 2 * F, TB, 2 * F))
            c.gemm([gemm_nn(ptr(self.dxhead), 2 * F, m.poff("dec_gauss_layer.mulayer.weight"), Hd,
                            ptr(self.dhA), Hd, TB, Hd, 2 * F)], mode)
            if late_disc:
                # Forked here, behind the same dependencies as the decoder's BPTT launch: the wavefront (high-priority
                # stream) takes its SMs first and the chain runs on the rest.  (Forked right behind the ELBO it held the
                # SMs while the BPTT launch was becoming resident: 127 vs 105 us for that launch.)
                self._disc_bwd(c, gflat)
            stack_bwd("dec", ptr(self.dhA), lambda l: None, defer=deferred if split_wgrad else None)
            wih_d = _lstm_names(pre["dec"], 0)[0]
            wg.append(gemm_tn(ptr(self.dgsum["dec", 0]), 4 * Hd, ptr(self.zcat), Z1 + Z2, g(wih_d), Z1 + Z2,
                              4 * Hd, Z1 + Z2, B))
            z1_from = dict(eps=self.eps1, head=self.z1head, roff=0, dgsum=ptr(self.dgsum["dec", 0]), NG=4 * Hd,
                           Wq=m.poff(wih_d), ld_wq=Z1 + Z2, Kq=Z1 + Z2, dzoff=0, beta=0)
        else:
            # reference behaviour (simple_fhvae.py:113-115): decoder and z samples get no gradient
            dec_names = [n for n in m._names if n.startswith(("pre_decoder", "dec_gauss_layer"))]
            lo = min(m._off[n] for n in dec_names)
            hi = max(m._off[n] + _prod(m._shape[n]) for n in dec_names)
            c.torch_op(lambda: (gflat[lo:hi].zero_(), self.dzcat.zero_()))
            assert all(lo <= m._off[n] < hi for n in dec_names)
            z1_from = {}
        # ---------------- z1 encoder
        Hz1 = self.H["z1"]
        head_bwd("z1", self.dz1head, Z1, "z1_gauss_layer.mulayer.weight", "z1_gauss_layer.mulayer.bias", **z1_from)
        while deferred:
            deferred.pop(0)()          # the decoder stack's layer-0 weight gradients: ready together with the z1 BPTT
        wih_z1 = _lstm_names(pre["z1"], 0)[0]

        def z1_extra():
            wgrad_tn(self.dg["z1", 0], 0, self.x_tm, 0, g(wih_z1), F + Z2, 4 * Hz1, F, TB)
            wg.append(gemm_tn(ptr(self.dgsum["z1", 0]), 4 * Hz1, ptr(self.zcat, Z1), Z1 + Z2, g(wih_z1, F), F + Z2,
                              4 * Hz1, Z2, B))
        stack_bwd("z1", None, lambda l: ptr(self.dhT["z1", l]), extra=z1_extra, defer=deferred if split_wgrad else None)
        # ---------------- z2 encoder
        Hz2 = self.H["z2"]
        # dz2_sample = (decoder part, already in dzcat[:, Z1:]) + dQ @ W_z ; dz2head also carries the side-2 chain's part
        c.join(2)
        head_bwd("z2", self.dz2head, Z2, "z2_gauss_layer.mulayer.weight", "z2_gauss_layer.mulayer.bias",
                 eps=self.eps2, head=self.z2head, roff=Z1, dgsum=ptr(self.dgsum["z1", 0]), NG=4 * Hz1, Wq=m.poff(wih_z1, F),
                 ld_wq=F + Z2, Kq=Z2, dzoff=Z1, beta=1)
        while deferred:
            deferred.pop(0)()          # the z1 stack's layer-0 weight gradients: beside the z2 BPTT
        wih_z2 = _lstm_names(pre["z2"], 0)[0]
        stack_bwd("z2", None, lambda l: ptr(self.dhT["z2", l]),
                  extra=lambda: wgrad_tn(self.dg["z2", 0], 0, self.x_tm, 0, g(wih_z2), F, 4 * Hz2, F, TB))
        wg.flush()
        return c


class _SimplePlan(_Plan):
    """Fully-connected variant (simple_fhvae.py:127-244): Linear+ReLU pre-encoders / pre-decoder."""

    def __init__(self, m: "SimpleFHVAE", B, T, F):
        super().__init__(m, B, T, F)
        f = self.f
        TF = T * F
        self.TF = TF
        self.a = {("z2", 0): f(B, m.z2_hus[0]), ("z2", 1): f(B, m.z2_hus[1]),
                  ("z1", 0): f(B, m.z1_hus[0]), ("z1", 1): f(B, m.z1_hus[1]),
                  ("dec", 0): f(B, m.x_hus[0]), ("dec", 1): f(B, m.x_hus[1])}
        self.xhead = f(B, 2 * TF)
        self._build_fwd()

    def px_views(self):
        xh = self.xhead
        return [xh[:, :self.TF].view(self.B, self.T, self.F), xh[:, self.TF:].view(self.B, self.T, self.F)]

    def _build_fwd(self):
        c, m, B, TF = self.fwd, self.m, self.B, self.TF
        Z1, Z2, mode = self.Z1, self.Z2, self.mode
        w = lambda n: m.poff(n + ".linear.weight")
        bia = lambda n: m.poff(n + ".linear.bias")
        h0, h1 = m.z2_hus
        x = ptr(self.x)
        c.gemm([gemm_nt(x, TF, w("z2_pre_encoder.fc1"), TF, ptr(self.a["z2", 0]), h0, B, h0, TF,
                        bias=bia("z2_pre_encoder.fc1"), relu=1)], mode)
        c.gemm([gemm_nt(ptr(self.a["z2", 0]), h0, w("z2_pre_encoder.fc2"), h0, ptr(self.a["z2", 1]), h1, B, h1,
                        h0, bias=bia("z2_pre_encoder.fc2"), relu=1)], mode)
        self._head_fwd(c, [(ptr(self.a["z2", 1]), h1, h1)], "z2_gauss_layer.mulayer.weight",
                       "z2_gauss_layer.mulayer.bias", self.z2head, Z2)
        c.add("fhvae_reparam_fwd", ptr(self.z2head), 2 * Z2, ptr(self.eps2), ptr(self.zcat, Z1), Z1 + Z2, B, Z2)
        h0, h1 = m.z1_hus
        ld1 = TF + Z2
        c.gemm([gemm_nt(x, TF, w("z1_pre_encoder.fc1"), ld1, ptr(self.a["z1", 0]), h0, B, h0, TF,
                        bias=bia("z1_pre_encoder.fc1"))], mode)
        c.gemm([gemm_nt(ptr(self.zcat, Z1), Z1 + Z2, m.poff("z1_pre_encoder.fc1.linear.weight", TF), ld1,
                        ptr(self.a["z1", 0]), h0, B, h0, Z2, beta=1.0, relu=1)], mode)
        c.gemm([gemm_nt(ptr(self.a["z1", 0]), h0, w("z1_pre_encoder.fc2"), h0, ptr(self.a["z1", 1]), h1, B, h1,
                        h0, bias=bia("z1_pre_encoder.fc2"), relu=1)], mode)
        self._head_fwd(c, [(ptr(self.a["z1", 1]), h1, h1)], "z1_gauss_layer.mulayer.weight",
                       "z1_gauss_layer.mulayer.bias", self.z1head, Z1)
        c.add("fhvae_reparam_fwd", ptr(self.z1head), 2 * Z1, ptr(self.eps1), ptr(self.zcat), Z1 + Z2, B, Z1)
        self.n_encode_calls = len(c.calls)
        h0, h1 = m.x_hus
        c.gemm([gemm_nt(ptr(self.zcat), Z1 + Z2, w("pre_decoder.fc1"), Z1 + Z2, ptr(self.a["dec", 0]), h0, B,
                        h0, Z1 + Z2, bias=bia("pre_decoder.fc1"), relu=1)], mode)
        c.gemm([gemm_nt(ptr(self.a["dec", 0]), h0, w("pre_decoder.fc2"), h0, ptr(self.a["dec", 1]), h1, B, h1,
                        h0, bias=bia("pre_decoder.fc2"), relu=1)], mode)
        c.gemm([gemm_nt(ptr(self.a["dec", 1]), h1, m.poff("dec_gauss_layer.mulayer.weight"), h1,
                        ptr(self.xhead), 2 * TF, B, 2 * TF, h1,
                        bias=m.poff("dec_gauss_layer.mulayer.bias"))], mode)
        self._tail_fwd(self.xhead, 2 * TF, self.F, TF)

    def _build_bwd(self, gflat) -> CallList:
        c, m, B, TF = CallList(), self.m, self.B, self.TF
        Z1, Z2, mode, f = self.Z1, self.Z2, self.mode, self.f
        g = lambda name, extra=0: ptr(gflat, m._off[name] + extra)
        if not hasattr(self, "dxhead"):
            self.dxhead = f(B, 2 * TF)
            self.da = {k: torch.zeros_like(v) for k, v in self.a.items()}
            self.dzcat = f(B, Z1 + Z2)
        self._tail_bwd(c, gflat, self.xhead, self.dxhead, 2 * TF, self.F, TF)
        c.join(2)              # dz2head / dmu2 carry the discriminative chain's part from here on
        wg, cs = [], []
        W = lambda n: n + ".linear.weight"
        Bn = lambda n: n + ".linear.bias"

        def linear_bwd(dy, N, xin, ldx, K, wname, bname, ldw=None, wcol=0, dx=None, lddx=None, beta=0.0,
                       bias=True):
            """dy (B,N) -> dW (N,K) [+ db], optional dx (B,K)."""
            ldw = ldw or K
            wg.append(gemm_tn(dy, N, xin, ldx, g(wname, wcol), ldw, N, K, B))
            if bias:
                cs.append(ColsumProblem(dy, g(bname), None, N, B, N))
            if dx is not None:
                c.gemm([gemm_nn(dy, N, m.poff(wname, wcol), ldw, dx, lddx, B, K, N, beta=beta)], mode)

        def relu_bwd(k):
            c.add("fhvae_relu_bwd", ptr(self.da[k]), ptr(self.a[k]), self.a[k].numel())

        # ---------------- decoder
        h0, h1 = m.x_hus
        if not m.detach_px:
            linear_bwd(ptr(self.dxhead), 2 * TF, ptr(self.a["dec", 1]), h1, h1, "dec_gauss_layer.mulayer.weight",
                       "dec_gauss_layer.mulayer.bias", dx=ptr(self.da["dec", 1]), lddx=h1)
            relu_bwd(("dec", 1))
            linear_bwd(ptr(self.da["dec", 1]), h1, ptr(self.a["dec", 0]), h0, h0, W("pre_decoder.fc2"),
                       Bn("pre_decoder.fc2"), dx=ptr(self.da["dec", 0]), lddx=h0)
            relu_bwd(("dec", 0))
            linear_bwd(ptr(self.da["dec", 0]), h0, ptr(self.zcat), Z1 + Z2, Z1 + Z2, W("pre_decoder.fc1"),
                       Bn("pre_decoder.fc1"), dx=ptr(self.dzcat), lddx=Z1 + Z2)
            c.add("fhvae_reparam_bwd", ptr(self.z1head), 2 * Z1, ptr(self.eps1), ptr(self.dzcat), Z1 + Z2,
                  ptr(self.dz1head), 2 * Z1, 1, B, Z1)
        else:
            dec_names = [n for n in m._names if n.startswith(("pre_decoder", "dec_gauss_layer"))]
            lo = min(m._off[n] for n in dec_names)
            hi = max(m._off[n] + _prod(m._shape[n]) for n in dec_names)
            assert all(lo <= m._off[n] < hi for n in dec_names)
            c.torch_op(lambda: (gflat[lo:hi].zero_(), self.dzcat.zero_()))
        # ---------------- z1 encoder
        h0, h1 = m.z1_hus
        ld1 = TF + Z2
        linear_bwd(ptr(self.dz1head), 2 * Z1, ptr(self.a["z1", 1]), h1, h1, "z1_gauss_layer.mulayer.weight",
                   "z1_gauss_layer.mulayer.bias", dx=ptr(self.da["z1", 1]), lddx=h1)
        relu_bwd(("z1", 1))
        linear_bwd(ptr(self.da["z1", 1]), h1, ptr(self.a["z1", 0]), h0, h0, W("z1_pre_encoder.fc2"),
                   Bn("z1_pre_encoder.fc2"), dx=ptr(self.da["z1", 0]), lddx=h0)
        relu_bwd(("z1", 0))
        linear_bwd(ptr(self.da["z1", 0]), h0, ptr(self.x), TF, TF, W("z1_pre_encoder.fc1"),
                   Bn("z1_pre_encoder.fc1"), ldw=ld1)
        linear_bwd(ptr(self.da["z1", 0]), h0, ptr(self.zcat, Z1), Z1 + Z2, Z2, W("z1_pre_encoder.fc1"), None,
                   ldw=ld1, wcol=TF, dx=ptr(self.dzcat, Z1), lddx=Z1 + Z2, beta=1.0, bias=False)
        c.add("fhvae_reparam_bwd", ptr(self.z2head), 2 * Z2, ptr(self.eps2), ptr(self.dzcat, Z1), Z1 + Z2,
              ptr(self.dz2head), 2 * Z2, 1, B, Z2)
        # ---------------- z2 encoder
        h0, h1 = m.z2_hus
        linear_bwd(ptr(self.dz2head), 2 * Z2, ptr(self.a["z2", 1]), h1, h1, "z2_gauss_layer.mulayer.weight",
                   "z2_gauss_layer.mulayer.bias", dx=ptr(self.da["z2", 1]), lddx=h1)
        relu_bwd(("z2", 1))
        linear_bwd(ptr(self.da["z2", 1]), h1, ptr(self.a["z2", 0]), h0, h0, W("z2_pre_encoder.fc2"),
                   Bn("z2_pre_encoder.fc2"), dx=ptr(self.da["z2", 0]), lddx=h0)
        relu_bwd(("z2", 0))
        linear_bwd(ptr(self.da["z2", 0]), h0, ptr(self.x), TF, TF, W("z2_pre_encoder.fc1"),
                   Bn("z2_pre_encoder.fc1"))
        c.gemm(wg, mode)
        c.colsum(cs)
        return c


# =====================================================================================
# public modules
# =====================================================================================
def _gauss_specs(prefix, in_dim, z):
    return [(f"{prefix}.mulayer.weight", (z, in_dim)), (f"{prefix}.logvar_layer.weight", (z, in_dim)),
            (f"{prefix}.mulayer.bias", (z,)), (f"{prefix}.logvar_layer.bias", (z,))]


def _check_z2(z2_dim):
    if z2_dim not in (8, 16, 32):
        raise ValueError(f"z2_dim={z2_dim}: the discriminative kernels (csrc/disc.cu) are built for z2_dim in "
                         "{8, 16, 32} (reference default 16, train_model.py:157-162; upstream 32)")


def _check_adjacent(off, shape, a, b):
    assert off[b] == off[a] + _prod(shape[a]), f"{a} and {b} must be adjacent in the flat buffer"


class SimpleFHVAE(_FHVAECore):
    """simple_fhvae.py:8-124.  Extra keyword-only arguments (all optional) configure what the
    reference leaves implicit: ``num_seqs`` (rows of the persistent mu2 table, Appendix A1),
    ``detach_px`` / ``prior_grad`` / ``ref_log_qy`` (reproduce the reference's gradient flow and its
    scalar +CE ``log_qy``: Appendix A2-A4), ``gemm_mode``, ``use_cuda_graphs``."""

    model = "simple_fhvae"

    def __init__(self, input_size, z1_hus=[128, 128], z2_hus=[128, 128], z1_dim=16, z2_dim=16,
                 x_hus=[128, 128], *, num_seqs=1000, init_std=1.0, detach_px=False, prior_grad=True,
                 ref_log_qy=False, gemm_mode=_lib.MODE_F32_SIMT, use_cuda_graphs=False):
        super().__init__()
        self.model = "simple_fhvae"
        self.pz1 = [0.0, np.float32(0.0)]
        self.pmu2 = [0.0, np.float32(PMU2_LOGVAR)]
        self.z1_hus, self.z2_hus, self.x_hus = _as_int_list(z1_hus), _as_int_list(z2_hus), _as_int_list(x_hus)
        self.z1_dim, self.z2_dim = int(z1_dim), int(z2_dim)
        self.input_size = int(input_size)
        self.detach_px, self.prior_grad, self.ref_log_qy = detach_px, prior_grad, ref_log_qy
        self.gemm_mode, self.use_cuda_graphs = gemm_mode, use_cuda_graphs
        assert len(self.z1_hus) == len(self.z2_hus) == len(self.x_hus) == 2, "two FC layers per block"
        assert self.z1_dim % 4 == 0 and self.z2_dim % 4 == 0
        _check_z2(self.z2_dim)
        I, Z1, Z2 = self.input_size, self.z1_dim, self.z2_dim
        # default nn.Linear init drawn in the reference's construction order (simple_fhvae.py:31-36)
        init, specs = {}, []

        def lin(name, i, o):
            l = nn.Linear(i, o)
            init[name + ".weight"], init[name + ".bias"] = l.weight, l.bias
            return [(name + ".weight", (o, i)), (name + ".bias", (o,))]

        def gauss(prefix, i, z):
            mu, lv = nn.Linear(i, z), nn.Linear(i, z)
            init[prefix + ".mulayer.weight"], init[prefix + ".mulayer.bias"] = mu.weight, mu.bias
            init[prefix + ".logvar_layer.weight"], init[prefix + ".logvar_layer.bias"] = lv.weight, lv.bias
            return _gauss_specs(prefix, i, z)

        specs += lin("z1_pre_encoder.fc1.linear", I + Z2, self.z1_hus[0])       # Appendix A6: z2 is concatenated
        specs += lin("z1_pre_encoder.fc2.linear", self.z1_hus[0], self.z1_hus[1])
        specs += lin("z2_pre_encoder.fc1.linear", I, self.z2_hus[0])
        specs += lin("z2_pre_encoder.fc2.linear", self.z2_hus[0], self.z2_hus[1])
        specs += gauss("z1_gauss_layer", self.z1_hus[1], Z1)
        specs += gauss("z2_gauss_layer", self.z2_hus[1], Z2)
        specs += lin("pre_decoder.fc1.linear", Z1 + Z2, self.x_hus[0])
        specs += lin("pre_decoder.fc2.linear", self.x_hus[0], self.x_hus[1])
        specs += gauss("dec_gauss_layer", self.x_hus[1], I)
        specs.append(("mu2_table", (int(num_seqs), Z2)))
        init["mu2_table"] = torch.empty(int(num_seqs), Z2).normal_(mean=0, std=init_std)   # :51
        self._init_flat(specs, init)
        for pfx in ("z1_gauss_layer", "z2_gauss_layer", "dec_gauss_layer"):
            _check_adjacent(self._off, self._shape, pfx + ".mulayer.weight", pfx + ".logvar_layer.weight")
            _check_adjacent(self._off, self._shape, pfx + ".mulayer.bias", pfx + ".logvar_layer.bias")

    def _make_plan(self, B, T, F):
        if T * F != self.input_size:
            raise ValueError(f"x is (B,{T},{F}) but input_size={self.input_size}")
        return _SimplePlan(self, B, T, F)


class FHVAE(_FHVAECore):
    """fhvae.py:4-14 signature; LSTM architecture per SURVEY.md Appendix B (the reference raises
    NotImplementedError).  ``input_size`` is seg_len * feat_dim as passed by train_model.py:398-402."""

    model = "fhvae"

    def __init__(self, input_size: int, z1_hus: list = [256, 256], z2_hus: list = [256, 256],
                 z1_dim: int = 32, z2_dim: int = 32, x_hus: list = [256, 256], *, seg_len=20,
                 num_seqs=1000, init_std=1.0, detach_px=False, prior_grad=True, ref_log_qy=False,
                 gemm_mode=_lib.MODE_F32_SIMT, use_cuda_graphs=False):
        super().__init__()
        self.model = "fhvae"
        self.pz1 = [0.0, np.float32(0.0)]
        self.pmu2 = [0.0, np.float32(PMU2_LOGVAR)]
        self.z1_hus, self.z2_hus, self.x_hus = _as_int_list(z1_hus), _as_int_list(z2_hus), _as_int_list(x_hus)
        self.z1_dim, self.z2_dim = int(z1_dim), int(z2_dim)
        self.input_size, self.seg_len = int(input_size), int(seg_len)
        if self.input_size % self.seg_len:
            raise ValueError("input_size must be seg_len * feat_dim")
        self.feat_dim = self.input_size // self.seg_len
        self.detach_px, self.prior_grad, self.ref_log_qy = detach_px, prior_grad, ref_log_qy
        self.gemm_mode, self.use_cuda_graphs = gemm_mode, use_cuda_graphs
        for hus in (self.z1_hus, self.z2_hus, self.x_hus):
            assert len(set(hus)) == 1 and hus[0] % 8 == 0, "one width per LSTM stack, multiple of 8"
        F, Z1, Z2 = self.feat_dim, self.z1_dim, self.z2_dim
        assert F % 4 == 0 and Z1 % 4 == 0 and Z2 % 4 == 0
        _check_z2(Z2)
        nets = [("z1_pre_encoder", "z1", F + Z2, self.z1_hus), ("z2_pre_encoder", "z2", F, self.z2_hus),
                ("pre_decoder", "dec", Z1 + Z2, self.x_hus)]
        init: Dict[str, torch.Tensor] = {}
        wspecs, bih, bhh, gspecs = [], [], [], {}

        def lstm(prefix, in_dim, hus):
            mod = nn.LSTM(in_dim, hus[0], num_layers=len(hus), batch_first=True)
            for n, p in mod.named_parameters():
                init[f"{prefix}.lstm.{n}"] = p
                tgt = wspecs if n.startswith("weight") else (bih if n.startswith("bias_ih") else bhh)
                tgt.append((f"{prefix}.lstm.{n}", tuple(p.shape)))

        def gauss(prefix, i, z):
            mu, lv = nn.Linear(i, z), nn.Linear(i, z)
            init[prefix + ".mulayer.weight"], init[prefix + ".mulayer.bias"] = mu.weight, mu.bias
            init[prefix + ".logvar_layer.weight"], init[prefix + ".logvar_layer.bias"] = lv.weight, lv.bias
            gspecs[prefix] = _gauss_specs(prefix, i, z)

        # construction order of the oracle (= reference order, simple_fhvae.py:31-36)
        lstm(*[nets[0][0], nets[0][2], nets[0][3]])
        lstm(*[nets[1][0], nets[1][2], nets[1][3]])
        gauss("z1_gauss_layer", sum(self.z1_hus), Z1)
        gauss("z2_gauss_layer", sum(self.z2_hus), Z2)
        lstm(*[nets[2][0], nets[2][2], nets[2][3]])
        gauss("dec_gauss_layer", self.x_hus[-1], F)
        # flat-buffer LAYOUT (independent of the registration order): [z2 encoder | z1 encoder + decoder | table].  The z2
        # encoder's gradients are the LAST ones a backward produces, so a data-parallel step can all-reduce the
        # [z1 + decoder + table] range while the z2 BPTT is still running (parallel.DataParallel, overlap).  Inside a
        # block: LSTM weights, all bias_ih, all bias_hh (one fused-bias launch per block), the Gaussian head.
        is_z2 = lambda ns: ns[0].startswith("z2_pre_encoder")
        blocks = [([w for w in wspecs if is_z2(w)], [b for b in bih if is_z2(b)], [b for b in bhh if is_z2(b)],
                   gspecs["z2_gauss_layer"]),
                  ([w for w in wspecs if not is_z2(w)], [b for b in bih if not is_z2(b)], [b for b in bhh if not is_z2(b)],
                   gspecs["z1_gauss_layer"] + gspecs["dec_gauss_layer"])]
        specs = [sp for blk in blocks for part in blk for sp in part]
        specs.append(("mu2_table", (int(num_seqs), Z2)))
        init["mu2_table"] = torch.empty(int(num_seqs), Z2).normal_(mean=0, std=init_std)
        self._init_flat(specs, init)
        self._z2_end = self._off[blocks[1][0][0][0]]            # flat offset where the z2-encoder block ends
        # fused-bias bookkeeping: per block the b_ih run and the b_hh run are contiguous and identically ordered
        self._bias_blocks, self._bias_off = [], {}
        short = {"z1_pre_encoder": "z1", "z2_pre_encoder": "z2", "pre_decoder": "dec"}
        boff = 0
        for _, bi, bh, _ in blocks:
            n = sum(_prod(sh) for _, sh in bi)
            assert self._off[bh[0][0]] == self._off[bi[0][0]] + n
            self._bias_blocks.append((bi[0][0], bh[0][0], n, boff))
            for name, _ in bi:
                prefix, _, leaf = name.split(".")
                self._bias_off[short[prefix], int(leaf.rsplit("l", 1)[1])] = boff + self._off[name] - self._off[bi[0][0]]
            boff += n
        self._bias_block_len = boff
        for pfx in ("z1_gauss_layer", "z2_gauss_layer", "dec_gauss_layer"):
            _check_adjacent(self._off, self._shape, pfx + ".mulayer.weight", pfx + ".logvar_layer.weight")
            _check_adjacent(self._off, self._shape, pfx + ".mulayer.bias", pfx + ".logvar_layer.bias")

    def _padded_batch(self, B, T):
        if B % 32 == 0 or os.environ.get("FHVAE_PAD_BATCH", "1") == "0":
            return B
        Bp = (B + 31) // 32 * 32
        sup = _lib.fn("fhvae_lstm_wave_supported")
        wave = any(len(h) == 2 and len(set(h)) == 1 and sup(T, Bp, int(h[0]), 2, self.gemm_mode)
                   for h in (self.z2_hus, self.z1_hus, self.x_hus))
        return Bp if wave else B

    def _make_plan(self, B, T, F):
        if T != self.seg_len or F != self.feat_dim:
            raise ValueError(f"x is (B,{T},{F}) but the model was built for seg_len={self.seg_len}, "
                             f"feat_dim={self.feat_dim}")
        return _FHVAEPlan(self, B, T, F)


def loss_function(lower_bound, log_qy, alpha=10.0):
    """train_model.py:243-251: ``-mean(lower_bound + alpha * log_qy)``.  Per-segment CUDA fp32 vectors take the
    one-launch path; anything else (the reference's scalar log_qy, other dtypes) the literal expression."""
    if (torch.is_tensor(log_qy) and lower_bound.is_cuda and log_qy.is_cuda and lower_bound.dtype == torch.float32
            and log_qy.dtype == torch.float32 and lower_bound.dim() == 1 and log_qy.shape == lower_bound.shape
            and lower_bound.is_contiguous() and log_qy.is_contiguous()):
        return _LossFn.apply(lower_bound, log_qy, alpha)
    return -1 * torch.mean(lower_bound + alpha * log_qy)
