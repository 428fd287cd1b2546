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


# ---------------------------------------------------------------------------------------------------------------
# Sharded table inside the train step: the exchange that model._Plan.run_train_step_sharded drives
# with CUDA kernels + NCCL, restated in torch over gloo (world 2) -- ownership routing, packet layout, rank-ordered
# LSE combine, sum_n p_bn m_n partial exchange, owner-local dense d table + owner scatter of the sparse rows --
# against autograd of the oracle's full-table discriminative term + KL(z2 || mu2) on the concatenated batch.
# ---------------------------------------------------------------------------------------------------------------
def _sharded_worker(rank, world, port, q):
    import math
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    N, Z, Bl = 23, 8, 6                                   # uneven shards (12 / 11 rows)
    Bg = Bl * world
    table = torch.randn(N, Z, dtype=torch.float64)
    zg = torch.randn(Bg, Z, dtype=torch.float64) * 0.5
    idxg = torch.randint(0, N, (Bg,))
    idxg[1] = idxg[0]; idxg[Bl] = idxg[0]                # duplicates inside and across ranks
    gg = torch.randn(Bg, dtype=torch.float64)             # dL/dlog_qy per segment
    # ---- reference: full table, concatenated batch, autograd
    t_ref, z_ref = table.clone().requires_grad_(True), zg.clone().requires_grad_(True)
    lq_ref = O.log_qy_per_segment(z_ref, t_ref, idxg)
    kl = -O.kld(z_ref, torch.zeros_like(z_ref), t_ref[idxg], O.PZ2_LOGVAR).sum(1)     # the sparse (row-gather) consumer
    ((gg * lq_ref).sum() + kl.sum()).backward()
    # ---- sharded: this rank holds rows rank::world and segments [rank*Bl, (rank+1)*Bl)
    shard = table[rank::world].clone()
    n_local = PP.shard_rows(N, rank, world)
    assert shard.shape[0] == n_local
    sl = slice(rank * Bl, (rank + 1) * Bl)
    z, idx, g = zg[sl], idxg[sl], gg[sl]
    # (1) packet all-gather: [z2_mu | idx as two 32-bit words | g | pad]
    pk = torch.zeros(Bl, Z + 4, dtype=torch.float64)
    pk[:, :Z] = z; pk[:, Z] = (idx & 0xffffffff).double(); pk[:, Z + 1] = (idx >> 32).double(); pk[:, Z + 2] = g
    parts = [torch.zeros_like(pk) for _ in range(world)]
    dist.all_gather(parts, pk)
    pkg = torch.cat(parts)
    z_all, g_all = pkg[:, :Z], pkg[:, Z + 2]
    idx_all = pkg[:, Z].long() | (pkg[:, Z + 1].long() << 32)
    assert torch.equal(idx_all, idxg)
    lidx = torch.where(PP.owner_of(idx_all, world) == rank, PP.local_row(idx_all, world), torch.full_like(idx_all, -1))
    # (2) owner-served mu2 rows: exactly one non-zero contributor per segment
    send = torch.zeros(Bg, Z, dtype=torch.float64)
    own = lidx >= 0
    send[own] = shard[lidx[own]]
    dist.all_reduce(send)
    mu2 = send[sl]
    assert torch.equal(mu2, table[idx])
    # (3) partial (max, sumexp) over the local rows for ALL global segments, all-gather, rank-ordered combine
    s_loc = O.disc_logits(z_all, shard)
    mx = s_loc.max(1).values
    part = torch.stack([mx, torch.exp(s_loc - mx[:, None]).sum(1)], -1)
    plist = [torch.zeros_like(part) for _ in range(world)]
    dist.all_gather(plist, part)
    lse_all = PP.combine_lse_partials(torch.stack(plist))
    tgt = -((z - mu2) ** 2).sum(1) / (2 * math.exp(O.PZ2_LOGVAR))
    lq = tgt - lse_all[sl]
    ok = torch.allclose(lq, lq_ref[sl].detach(), rtol=1e-12, atol=1e-12)
    # (4) backward: p_bn over local rows; dense d table owner-local; sum_n p_bn m_n partials to the segment's rank
    p = torch.exp(s_loc - lse_all[:, None])                                  # (Bg, n_local)
    inv_s2 = 1.0 / math.exp(O.PZ2_LOGVAR)
    dshard = -(g_all[:, None, None] * p[:, :, None] * (z_all[:, None, :] - shard[None]) * inv_s2).sum(0)
    sumpm = p @ shard                                                        # (Bg, Z) partial over local rows
    slist = [torch.zeros_like(sumpm) for _ in range(world)]
    dist.all_gather(slist, sumpm)
    sp = sum(s_[sl] for s_ in slist)                                         # fixed rank order
    dz = g[:, None] * inv_s2 * (mu2 - sp) + (-(z - mu2) * inv_s2)            # disc part + KL(z2 || mu2) part
    ok = ok and torch.allclose(dz, z_ref.grad[sl], rtol=1e-10, atol=1e-12)
    # (5) sparse rows (target part of log q + KL part): all-gather, owner scatters in ascending global segment order
    dmu2 = g[:, None] * inv_s2 * (z - mu2) + (z - mu2) * inv_s2
    dlist = [torch.zeros_like(dmu2) for _ in range(world)]
    dist.all_gather(dlist, dmu2)
    dmu2_all = torch.cat(dlist)
    touched = []
    for b in range(Bg):
        if lidx[b] >= 0:
            if int(lidx[b]) not in [int(lidx[j]) for j in range(b) if lidx[j] >= 0]:
                touched.append(b)
            dshard[lidx[b]] += dmu2_all[b]
    ok = ok and torch.allclose(dshard, t_ref.grad[rank::world], rtol=1e-10, atol=1e-12)
    first = [b for b in range(Bg) if int(idxg[b]) % world == rank and int(idxg[b]) not in idxg[:b].tolist()]
    ok = ok and touched == first
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_table_exchange_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_alloc_and_dp_argument_checks():
    assert PP.shard_alloc_rows(203, 2) == 102 and PP.shard_alloc_rows(5000, 8) == 625 and PP.shard_alloc_rows(5, 8) == 1
    assert [PP.shard_rows(5, r, 8) for r in range(8)] == [1, 1, 1, 1, 1, 0, 0, 0]
