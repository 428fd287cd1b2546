"""CPU, gloo, world_size 2: host-side logic of the multi-GPU path (ownership / routing of table rows,
rank-ordered LSE combine, flat-gradient all-reduce wiring)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fhvae_oracle as O
from pytorch_scalablefhvae_b200 import parallel as PP


def test_ownership_partitions_rows_exactly():
    N, W = 280000, 8
    u = torch.arange(N)
    own, loc = PP.owner_of(u, W), PP.local_row(u, W)
    assert sum(PP.shard_rows(N, r, W) for r in range(W)) == N
    for r in range(W):
        mine = (own == r)
        assert int(mine.sum()) == PP.shard_rows(N, r, W)
        assert torch.equal(loc[mine], torch.arange(int(mine.sum())))           # dense local numbering
        assert torch.equal(loc[mine] * W + r, u[mine])                          # invertible
    pos, rows = PP.route_to_owners(torch.tensor([5, 8, 16, 7, 0]), 8, 0)
    assert pos.tolist() == [1, 2, 4] and rows.tolist() == [1, 2, 0]


def test_sharded_lse_equals_full_lse():
    g = torch.Generator().manual_seed(0)
    B, N, Z, W = 9, 50, 8, 4
    z, table = torch.randn(B, Z, generator=g), torch.randn(N, Z, generator=g)
    logits = O.disc_logits(z, table)
    parts = []
    for r in range(W):
        lg = logits[:, r::W]
        m = lg.max(dim=1).values
        parts.append(torch.stack([m, torch.exp(lg - m[:, None]).sum(1)], -1))
    parts.append(torch.stack([torch.full((B,), float("-inf")), torch.zeros(B)], -1))   # an empty shard
    torch.testing.assert_close(PP.combine_lse_partials(torch.stack(parts)), torch.logsumexp(logits, 1))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class _M:            # minimal stand-in exposing what DataParallel touches
        def __init__(self):
            self.flat = torch.full((10,), float(rank + 1))
        def _ensure_flat(self):
            return self.flat

    class _Opt:
        grad_scale = 1.0

    m, o = _M(), _Opt()
    dp = PP.DataParallel(m, o)
    ok = (o.grad_scale == 0.5) and bool((m.flat == 1.0).all())                 # broadcast from rank 0
    g = torch.full((6,), float(rank + 1))
    dp.allreduce_(g)
    ok = ok and bool((g == 3.0).all())
    ok = ok and float(dp.global_mean(torch.tensor(float(rank)))) == 0.5
    # the same utterance list routes to disjoint owners that together cover it
    utt = torch.tensor([3, 10, 11, 4, 280001])
    pos, rows = PP.route_to_owners(utt, world, rank)
    cnt = torch.zeros(len(utt)); cnt[pos] = 1
    dist.all_reduce(cnt)
    ok = ok and bool((cnt == 1).all()) and torch.equal(rows * world + rank, utt[pos])
    q.put((rank, ok))
    dist.destroy_process_group()


def test_data_parallel_wiring_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert sorted(res) == [(0, True), (1, True)]
