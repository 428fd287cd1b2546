"""Data-parallel training over the GPUs of one box: one process per GPU (torchrun), segments are the
sharded unit.  The reference has no multi-device code at all.

What is exchanged per step
* dense parameters are replicated; their flat fp32 gradient buffer (10.96 MB for config 1) is summed
  with ONE NCCL all-reduce over NVLink/NVSwitch and scaled by 1/world inside the Adam kernel;
* the mu2 table is either replicated (its gradient rides in the same flat all-reduce) or SHARDED by row id
  (row u lives on rank u mod W): see DataParallel.  The same ownership rule shards the 280k-row master table of
  the hierarchical-sampling mode (hierarchical.py).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


# ------------------------------------------------------------------ ownership of master-table rows
def owner_of(utt: torch.Tensor, world: int) -> torch.Tensor:
    """Rank that owns master-table row `utt` (int64 tensor): utt mod world."""
    return torch.remainder(utt, world)


def local_row(utt: torch.Tensor, world: int) -> torch.Tensor:
    """Row index inside the owner's shard."""
    return torch.div(utt, world, rounding_mode="floor")


def shard_rows(num_rows: int, rank: int, world: int) -> int:
    """Number of master rows held by `rank`."""
    return (num_rows - rank + world - 1) // world


def route_to_owners(utt: torch.Tensor, world: int, rank: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """For a list of utterance ids (same on every rank) return (positions, local rows) of the ids this
    rank owns -- exact int64 arithmetic; positions index into `utt`."""
    mine = (owner_of(utt, world) == rank).nonzero(as_tuple=True)[0]
    return mine, local_row(utt[mine], world)


# ------------------------------------------------------------------ log-sum-exp over sharded rows
def combine_lse_partials(parts: torch.Tensor) -> torch.Tensor:
    """parts (P, B, 2) = (max, sum exp(s - max)) per shard/split -> logsumexp (B,).  Fixed, rank-ordered
    combine so that every rank computes bit-identical values."""
    m = parts[..., 0]
    M = m.max(dim=0).values
    w = torch.where(torch.isinf(m) & (m < 0), torch.zeros_like(m), torch.exp(m - M))
    return M + torch.log((parts[..., 1] * w).sum(dim=0))


# ------------------------------------------------------------------ the DP wrapper
def shard_alloc_rows(num_rows: int, world: int) -> int:
    """Rows every rank allocates for its shard (the largest shard; ranks with one row less leave the last unused)."""
    return (int(num_rows) + world - 1) // world


class DataParallel:
    """Wraps (model, FusedAdam).  ``train_step`` = local fwd+bwd on this rank's segments, all-reduce of the dense
    gradients, fused Adam with grad_scale = 1/world (loss = mean over the GLOBAL batch).

    table="replicated": every rank holds the whole mu2 table; its gradient rides in the same flat all-reduce
        (exact: the softmax over all rows couples every row to every segment).
    table="sharded" (north star): row u lives on rank ``u mod W`` at local row ``u // W``; the model is built with
        ``num_seqs = shard_alloc_rows(num_rows, W)`` and ``mu_idx`` passed to ``train_step`` are GLOBAL row ids.
        Per step (model._Plan.run_train_step_sharded): all-gather (z2_mu, idx, g) -> every rank scores all B_global
        segments against ITS rows -> all-gather of (max, sumexp) partials, rank-ordered combine -> all-to-all of the
        sum_n p_bn m_n partials for dz2_mu; the dense softmax part of d table never leaves its owner, the sparse
        KL / prior / target row gradients are all-gathered and scatter-reduced by the owner in ascending global
        segment order; Adam updates owned rows only.  With world == 1 (no process group) the collectives are
        copies, so the whole sharded path also runs -- and is tested -- on one GPU."""

    def __init__(self, model, optimizer, group: Optional[dist.ProcessGroup] = None, table: str = "replicated",
                 num_rows: Optional[int] = None, overlap: bool = False):
        self.model, self.optimizer, self.group = model, optimizer, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if table not in ("replicated", "sharded"):
            raise ValueError("table must be 'replicated' or 'sharded'")
        self.table = table
        # replicated mode, optional: all-reduce [z1 + decoder + table] beside the z2 BPTT (3 graphs + 2 all-reduces).
        # Measured on 2 x B200: 1.020 ms/step vs 0.987 ms with ONE all-reduce between backward and Adam (the fixed
        # latency of a second NCCL launch + the graph cuts cost more than the overlap hides), so it is off by default.
        self.overlap = bool(overlap)
        optimizer.grad_scale = 1.0 / self.world
        if table == "sharded":
            if num_rows is None:
                raise ValueError("table='sharded' needs num_rows (rows of the whole table)")
            self.num_rows = int(num_rows)
            self.n_local = shard_rows(self.num_rows, self.rank, self.world)
            alloc = shard_alloc_rows(self.num_rows, self.world)
            if model.mu2_table.shape[0] != alloc:
                raise ValueError(f"sharded table of {num_rows} rows over {self.world} ranks: build the model with "
                                 f"num_seqs={alloc} (shard_alloc_rows), got {model.mu2_table.shape[0]}")
        if self.world > 1:
            self.broadcast_parameters()

    # ---- parameters
    def broadcast_parameters(self):
        flat = self.model._ensure_flat()
        if self.table == "sharded":                 # shards are per-rank state: only the dense prefix is replicated
            flat = flat[:self.model._off["mu2_table"]]
        dist.broadcast(flat, src=0, group=self.group)

    @torch.no_grad()
    def shard_table_(self, full_table: torch.Tensor):
        """Load this rank's rows (rank, rank+W, ...) of a full (num_rows, Z) table into the model's shard."""
        assert self.table == "sharded" and full_table.shape[0] == self.num_rows
        t = self.model.mu2_table
        t.zero_()
        t[:self.n_local].copy_(full_table[self.rank::self.world].to(t.device))

    @torch.no_grad()
    def gather_table(self) -> torch.Tensor:
        """The full (num_rows, Z) table assembled from every rank's shard (identical on all ranks)."""
        assert self.table == "sharded"
        t = self.model.mu2_table.detach()
        alloc, Z = t.shape
        allsh = torch.empty(self.world, alloc, Z, device=t.device)
        self.all_gather(allsh.view(self.world * alloc, Z), t.contiguous())
        full = torch.empty(self.num_rows, Z, device=t.device)
        for r in range(self.world):
            n = shard_rows(self.num_rows, r, self.world)
            full[r::self.world] = allsh[r, :n]
        return full

    # ---- collectives (copies when there is no process group: the sharded path stays testable on one GPU)
    def allreduce_(self, gflat: torch.Tensor):
        if self.world > 1:
            dist.all_reduce(gflat, group=self.group)

    def all_gather(self, out: torch.Tensor, inp: torch.Tensor):
        if self.world > 1:
            dist.all_gather_into_tensor(out, inp, group=self.group)
        else:
            out.view(-1).copy_(inp.reshape(-1))

    def reduce_scatter(self, out: torch.Tensor, inp: torch.Tensor):
        if self.world > 1:
            dist.reduce_scatter_tensor(out, inp, group=self.group)
        else:
            out.view(-1).copy_(inp.reshape(-1))

    def all_to_all(self, out: torch.Tensor, inp: torch.Tensor):
        if self.world > 1:
            dist.all_to_all_single(out, inp, group=self.group)
        else:
            out.view(-1).copy_(inp.reshape(-1))

    def train_step(self, x, mu_idx, num_segs, alpha: float = 10.0, eps=None):
        if self.table == "sharded":
            return self.model.train_step(x, mu_idx, num_segs, self.optimizer, alpha, eps=eps, shard=self)
        return self.model.train_step(x, mu_idx, num_segs, self.optimizer, alpha, eps=eps,
                                     allreduce=self.allreduce_ if self.world > 1 else None,
                                     overlap=self if (self.overlap and self.world > 1) else None)

    def global_mean(self, local_scalar: torch.Tensor) -> torch.Tensor:
        t = local_scalar.detach().clone()
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
            t /= self.world
        return t
