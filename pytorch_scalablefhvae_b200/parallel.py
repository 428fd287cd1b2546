"""Data-parallel training over the GPUs of one box: one process per GPU (torchrun), segments are the
sharded unit.  The reference has no multi-device code at all.

What is exchanged per step
* dense parameters are replicated; their flat fp32 gradient buffer (10.96 MB for config 1) is summed
  with ONE NCCL all-reduce over NVLink/NVSwitch and scaled by 1/world inside the Adam kernel;
* the mu2 table: this round keeps the *active* table (N rows, 128 KB at N=1000 / 640 KB at K=5000)
  replicated and all-reduces its gradient in the same flat buffer, which is exact (the softmax over all
  rows couples every row to every segment).  Ownership helpers for the sharded
  master table (row u lives on rank u mod W) are below and are what the hierarchical-sampling cache
  refresh uses to route rows to their owner.
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
class DataParallel:
    """Wraps (model, FusedAdam): `train_step` = local fwd+bwd, one all-reduce of the flat gradient
    buffer, fused Adam with grad_scale = 1/world (loss = mean over the GLOBAL batch)."""

    def __init__(self, model, optimizer, group: Optional[dist.ProcessGroup] = None):
        self.model, self.optimizer, self.group = model, optimizer, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        optimizer.grad_scale = 1.0 / self.world
        if self.world > 1:
            self.broadcast_parameters()

    def broadcast_parameters(self):
        flat = self.model._ensure_flat()
        dist.broadcast(flat, src=0, group=self.group)

    def allreduce_(self, gflat: torch.Tensor):
        if self.world > 1:
            dist.all_reduce(gflat, group=self.group)

    def train_step(self, x, mu_idx, num_segs, alpha: float = 10.0, eps=None):
        return self.model.train_step(x, mu_idx, num_segs, self.optimizer, alpha, eps=eps,
                                     allreduce=self.allreduce_ if self.world > 1 else None)

    def global_mean(self, local_scalar: torch.Tensor) -> torch.Tensor:
        t = local_scalar.detach().clone()
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
            t /= self.world
        return t
