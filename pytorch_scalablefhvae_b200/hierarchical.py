"""Hierarchical sampling (BASELINE config 3; the broken skeleton at train_model.py:424-436).

The master mu2 table (e.g. 280,000 utterances) is SHARDED by utterance id: row u lives on rank u mod W at
local row u // W.  Training touches only a cache of K sampled utterances (K = 5000,
train_model.py:209-214) whose rows are the model's `mu2_table` parameter (replicated, local label =
position in the sampled list, train_model.py:436).  Per round:
  sample()      bit-exact `np.random.choice(seqlist, K, replace=False)` on the legacy numpy RNG
  fetch()       cache <- owners' rows           (exact int64 routing; one all-reduce of K x Z floats)
  refresh()     cache <- MAP estimate with the current encoder, utils.estimate_mu2_dict (utils.py:45-60)
  write_back()  owners' rows <- cache           (sparse row update, owner-local, nothing else moves)
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .inference import R_MU2
from .parallel import DataParallel, route_to_owners, shard_alloc_rows, shard_rows
from .plan import current_stream_ptr, ptr


def sample_sequences(seqlist: Sequence, k: int, seed: int) -> np.ndarray:
    """train_model.py:426-428 with the global legacy RNG seeded: identical set AND order on every rank."""
    return np.random.RandomState(seed).choice(np.asarray(seqlist), k, replace=False)


class ShardedMu2Table:
    def __init__(self, num_utts: int, z2_dim: int, device, init_std: float = 1.0, seed: int = 99,
                 group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.num_utts, self.z2_dim, self.device = int(num_utts), int(z2_dim), device
        n_local = shard_rows(self.num_utts, self.rank, self.world)
        g = torch.Generator().manual_seed(seed + self.rank)
        self.shard = (torch.randn(max(n_local, 1), z2_dim, generator=g) * init_std).to(device)   # simple_fhvae.py:51

    def _route(self, utts: torch.Tensor):
        pos, rows = route_to_owners(utts.cpu(), self.world, self.rank)
        return pos.to(self.device), rows.to(self.device)

    def fetch(self, utts: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(K, Z) rows of the listed utterances, identical on every rank."""
        K = utts.numel()
        cache = torch.zeros(K, self.z2_dim, device=self.device) if out is None else out.zero_()
        pos, rows = self._route(utts)
        if pos.numel():
            _lib.check(_lib.fn("fhvae_rows_copy")(ptr(self.shard), ptr(rows), ptr(cache), ptr(pos), pos.numel(),
                                                  self.z2_dim, current_stream_ptr()), "fhvae_rows_copy")
        if self.world > 1:
            dist.all_reduce(cache, group=self.group)          # every row has exactly one non-zero contributor
        return cache

    def write_back(self, utts: torch.Tensor, cache: torch.Tensor):
        """Owner-local sparse row update: shard[u // W] = cache[position of u] for the u this rank owns."""
        pos, rows = self._route(utts)
        if pos.numel():
            _lib.check(_lib.fn("fhvae_rows_copy")(ptr(cache), ptr(pos), ptr(self.shard), ptr(rows), pos.numel(),
                                                  self.z2_dim, current_stream_ptr()), "fhvae_rows_copy")

    @torch.no_grad()
    def refresh(self, model, x_batches, label_batches) -> torch.Tensor:
        """MAP re-estimate of the K cache rows with the current encoder.  `x_batches` yields (B,T,F) CUDA
        segment batches of THIS rank's share, `label_batches` their local labels (B,) int64 in [0,K).
        Writes the result into model.mu2_table and returns it."""
        K, Z = model.mu2_table.shape
        zsum, cnt = torch.zeros(K, Z, device=self.device), torch.zeros(K, device=self.device)
        acc = _lib.fn("fhvae_mu2_accumulate")
        err = torch.zeros(1, dtype=torch.int32, device=self.device)
        for x, lab in zip(x_batches, label_batches):
            enc = model.encode(x)
            lab = lab.to(self.device)
            _lib.check(acc(ptr(enc["z2_mu"]), 2 * Z, ptr(lab), ptr(zsum), ptr(cnt), x.shape[0], Z, K, ptr(err),
                           current_stream_ptr()), "fhvae_mu2_accumulate")
        if int(err):                                   # round-level operation: one host read is fine here
            raise IndexError("refresh: a label is outside [0, K) of the sampled cache")
        if self.world > 1:
            dist.all_reduce(zsum, group=self.group)
            dist.all_reduce(cnt, group=self.group)
        table = model.mu2_table.data
        _lib.check(_lib.fn("fhvae_mu2_estimate_finish")(ptr(zsum), ptr(cnt), ptr(table), R_MU2, K, Z,
                                                        current_stream_ptr()), "fhvae_mu2_estimate_finish")
        return table



class HierarchicalTrainer:
    """One ROUND of hierarchical sampling end to end (BASELINE config 3; what train_model.py:424-436 sketches):

        sample K utterances (bit-exact np.random.choice) -> fetch their rows from the owners of the sharded master
        table -> [refresh: MAP re-estimate with the current encoder] -> `steps` data-parallel train steps on segments
        of those utterances, the ACTIVE K-row table sharded by label (label l on rank l mod W) inside the train step
        -> write the trained rows back to their owners.

    ``model`` must be built with ``num_seqs = shard_alloc_rows(K, world)``; ``master`` is a ShardedMu2Table over all
    utterances.  Labels handed to ``train_step`` are positions in the sampled list (train_model.py:436)."""

    def __init__(self, model, optimizer, master: ShardedMu2Table, k: int, group: Optional[dist.ProcessGroup] = None):
        self.model, self.optimizer, self.master, self.k = model, optimizer, master, int(k)
        self.dp = DataParallel(model, optimizer, group=group, table="sharded", num_rows=self.k)
        self.utts: Optional[torch.Tensor] = None
        self.rounds = 0

    def begin_round(self, seed: int, refresh=None) -> torch.Tensor:
        """Sample + fetch (+ refresh).  Returns the sampled utterance ids (K,) int64, identical on every rank.
        ``refresh``: optional (x_batches, label_batches) of this rank's segments for the MAP re-estimate."""
        ids = np.arange(self.master.num_utts)
        sel = np.random.RandomState(seed).choice(ids, self.k, replace=False)      # train_model.py:426-428
        self.utts = torch.from_numpy(sel.astype(np.int64))
        cache = self.master.fetch(self.utts)                                      # (K, Z) identical on every rank
        if refresh is not None:
            cache = self._refresh(cache, *refresh)
        self.dp.shard_table_(cache)                                               # active rows label::W on this rank
        self.optimizer.reset_state(self.model, "mu2_table")
        self.rounds += 1
        return self.utts

    @torch.no_grad()
    def _refresh(self, cache, x_batches, label_batches):
        K, Z = cache.shape
        dev = cache.device
        zsum, cnt = torch.zeros(K, Z, device=dev), torch.zeros(K, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        acc = _lib.fn("fhvae_mu2_accumulate")
        for x, lab in zip(x_batches, label_batches):
            enc = self.model.encode(x)
            lab = lab.to(dev)
            _lib.check(acc(ptr(enc["z2_mu"]), 2 * Z, ptr(lab), ptr(zsum), ptr(cnt), x.shape[0], Z, K, ptr(err),
                           current_stream_ptr()), "fhvae_mu2_accumulate")
        if int(err):
            raise IndexError("refresh: a label is outside [0, K) of the sampled cache")
        if self.dp.world > 1:
            dist.all_reduce(zsum, group=self.dp.group)
            dist.all_reduce(cnt, group=self.dp.group)
        out = cache.clone()
        _lib.check(_lib.fn("fhvae_mu2_estimate_finish")(ptr(zsum), ptr(cnt), ptr(out), R_MU2, K, Z,
                                                        current_stream_ptr()), "fhvae_mu2_estimate_finish")
        return out

    def train_step(self, x, labels, num_segs, alpha: float = 10.0, eps=None):
        """labels (B,) int64 in [0, K): positions of the segments' utterances in the sampled list."""
        return self.dp.train_step(x, labels, num_segs, alpha, eps=eps)

    def end_round(self):
        """Trained rows -> their owners in the master table (owner-local sparse row update)."""
        cache = self.dp.gather_table()
        self.master.write_back(self.utts, cache)
        return cache
