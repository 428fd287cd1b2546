"""Run under torchrun on N GPUs: N-rank data-parallel training must equal single-GPU training on the
concatenated batch (same parameters after two Adam steps).  Exit code 0 = pass."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.distributed as dist

import pytorch_scalablefhvae_b200 as P
from pytorch_scalablefhvae_b200.parallel import DataParallel
from util import synth_batch


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    Bl, T, F, N, H, Z = 64, 20, 80, 200, 256, 32
    mode = P.MODE_BF16X3
    args = (T * F, [H, H], [H, H], Z, Z, [H, H])
    torch.manual_seed(0)
    m = P.FHVAE(*args, seg_len=T, num_seqs=N, gemm_mode=mode, use_cuda_graphs=True).to(dev)   # overlapped all-reduce path
    opt = P.FusedAdam(m.parameters(), lr=1e-3, betas=(0.95, 0.999))
    dp = DataParallel(m, opt, overlap=True)
    torch.manual_seed(0)
    ref = P.FHVAE(*args, seg_len=T, num_seqs=N, gemm_mode=mode).to(dev)
    ropt = P.FusedAdam(ref.parameters(), lr=1e-3, betas=(0.95, 0.999))
    worst = 0.0
    for step in range(3):
        x, idx, nsegs = synth_batch(Bl * world, T, F, N, seed=50 + step)
        g = torch.Generator().manual_seed(step)
        eps = {"z1": torch.randn(Bl * world, Z, generator=g), "z2": torch.randn(Bl * world, Z, generator=g)}
        sl = slice(rank * Bl, (rank + 1) * Bl)
        l_dp = dp.train_step(x[sl].to(dev), idx[sl], nsegs[sl], 10.0, eps={k: v[sl] for k, v in eps.items()})
        l_ref = ref.train_step(x.to(dev), idx, nsegs, ropt, 10.0, eps=eps)
        worst = max(worst, abs(float(dp.global_mean(l_dp)) - float(l_ref)) / abs(float(l_ref)))
    for (k, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        d = float((p - q).abs().max())
        assert d <= 2 * 3 * 1e-3, (k, d)                                       # bounded (Adam sign-like at |g|~0)
        frac = float(((p - q).abs() <= 1e-4 * float(q.abs().max())).float().mean())
        assert frac >= 0.999, (k, frac)
    assert worst < 1e-4, worst
    # replicas stay bit-identical across ranks
    flat = m._flat.clone()
    dist.broadcast(flat, src=0)
    assert torch.equal(flat, m._flat)
    # ---- sharded master table: rows routed to / fetched from their owner (utt mod world), exact
    from pytorch_scalablefhvae_b200.parallel import shard_rows
    Nm, K = 1003, 64
    table = P.ShardedMu2Table(Nm, Z, dev, seed=3)
    full = torch.zeros(Nm, Z)
    for r in range(world):                       # what every rank's shard holds, regenerated from the seeds
        g = torch.Generator().manual_seed(3 + r)
        full[r::world] = torch.randn(max(shard_rows(Nm, r, world), 1), Z, generator=g)[:shard_rows(Nm, r, world)]
    utts = torch.from_numpy(__import__("numpy").random.RandomState(11).permutation(Nm)[:K].astype("int64"))
    cache = table.fetch(utts)
    assert torch.equal(cache.cpu(), full[utts]), "fetch must return the owners' rows exactly"
    table.write_back(utts, cache * 2.0)
    again = table.fetch(utts)
    assert torch.equal(again.cpu(), full[utts] * 2.0)
    other = torch.tensor([u for u in range(Nm) if u not in set(utts.tolist())][:50])
    assert torch.equal(table.fetch(other).cpu(), full[other]), "rows outside the sample stay untouched"
    # ---- mu2 table SHARDED inside the train step (row u on rank u mod W): W-rank step == 1-GPU full-table step on
    # the concatenated batch -- losses / dense parameters 1e-4, assembled table rows 1e-4, touched-row sets bit-exact
    from pytorch_scalablefhvae_b200.parallel import shard_alloc_rows, owner_of
    Nrows = 203                                                    # uneven shards
    torch.manual_seed(1)
    full = P.FHVAE(*args, seg_len=T, num_seqs=Nrows, gemm_mode=mode, use_cuda_graphs=True).to(dev)
    fopt = P.FusedAdam(full.parameters(), lr=1e-3, betas=(0.95, 0.999))
    torch.manual_seed(1)
    sm = P.FHVAE(*args, seg_len=T, num_seqs=shard_alloc_rows(Nrows, world), gemm_mode=mode, use_cuda_graphs=True).to(dev)
    sopt = P.FusedAdam(sm.parameters(), lr=1e-3, betas=(0.95, 0.999))
    sdp = DataParallel(sm, sopt, table="sharded", num_rows=Nrows)
    sdp.shard_table_(full.mu2_table.detach())
    worst_s = 0.0
    for step in range(3):
        x, idx, nsegs = synth_batch(Bl * world, T, F, Nrows, seed=70 + step)
        g = torch.Generator().manual_seed(10 + step)
        eps = {"z1": torch.randn(Bl * world, Z, generator=g), "z2": torch.randn(Bl * world, Z, generator=g)}
        sl = slice(rank * Bl, (rank + 1) * Bl)
        l_s = sdp.train_step(x[sl].to(dev), idx[sl].to(dev), nsegs[sl].to(dev), 10.0, eps={k: v[sl] for k, v in eps.items()})
        l_f = full.train_step(x.to(dev), idx.to(dev), nsegs.to(dev), fopt, 10.0, eps=eps)
        worst_s = max(worst_s, abs(float(sdp.global_mean(l_s)) - float(l_f)) / abs(float(l_f)))
        # touched rows: this rank's flags == first occurrences (in global segment order) of the rows it owns
        first, seen = torch.zeros(Bl * world, dtype=torch.int32), set()
        for b, r in enumerate(idx.tolist()):
            if r not in seen and r % world == rank:
                first[b] = 1
            seen.add(r)
        assert torch.equal(sm._plan(Bl, T, F).touched_global.cpu(), first), "touched-row set"
    assert worst_s < 1e-4, worst_s
    sm.check_flags()
    tab = sdp.gather_table()
    d = (tab - full.mu2_table.detach()).abs()
    assert float(d.max()) <= 2 * 3 * 1e-3 and float((d <= 1e-4 * float(full.mu2_table.abs().max())).float().mean()) >= 0.999
    for (k, p), (_, q) in zip(sm.named_parameters(), full.named_parameters()):
        if k == "mu2_table":
            continue
        d = float((p - q).abs().max())
        assert d <= 2 * 3 * 1e-3, (k, d)
        frac = float(((p - q).abs() <= 1e-4 * float(q.abs().max())).float().mean())
        assert frac >= 0.999, (k, frac)
    dense = sm._flat[:sm._off["mu2_table"]].clone()
    dist.broadcast(dense, src=0)
    assert torch.equal(dense, sm._flat[:sm._off["mu2_table"]]), "dense replicas must stay bit-identical"
    # ---- one hierarchical round (BASELINE config 3) over W ranks == single-GPU training on the fetched rows
    Nm, K = 1003, 64
    torch.manual_seed(2)
    hm = P.FHVAE(*args, seg_len=T, num_seqs=shard_alloc_rows(K, world), gemm_mode=mode, use_cuda_graphs=True).to(dev)
    hopt = P.FusedAdam(hm.parameters(), lr=1e-3, betas=(0.95, 0.999))
    torch.manual_seed(2)
    hr = P.FHVAE(*args, seg_len=T, num_seqs=K, gemm_mode=mode, use_cuda_graphs=True).to(dev)
    hropt = P.FusedAdam(hr.parameters(), lr=1e-3, betas=(0.95, 0.999))
    master = P.ShardedMu2Table(Nm, Z, dev, seed=5)
    tr = P.HierarchicalTrainer(hm, hopt, master, K)
    utts = tr.begin_round(seed=21)
    assert utts.tolist() == __import__("numpy").random.RandomState(21).choice(__import__("numpy").arange(Nm), K, replace=False).tolist()
    cache0 = master.fetch(utts)
    with torch.no_grad():
        hr.mu2_table.copy_(cache0)
    worst_h = 0.0
    for step in range(2):
        x, lab, nsegs = synth_batch(Bl * world, T, F, K, seed=90 + step)
        g = torch.Generator().manual_seed(30 + step)
        eps = {"z1": torch.randn(Bl * world, Z, generator=g), "z2": torch.randn(Bl * world, Z, generator=g)}
        sl = slice(rank * Bl, (rank + 1) * Bl)
        l_h = tr.train_step(x[sl].to(dev), lab[sl].to(dev), nsegs[sl].to(dev), 10.0, eps={k: v[sl] for k, v in eps.items()})
        l_r = hr.train_step(x.to(dev), lab.to(dev), nsegs.to(dev), hropt, 10.0, eps=eps)
        worst_h = max(worst_h, abs(float(tr.dp.global_mean(l_h)) - float(l_r)) / abs(float(l_r)))
    assert worst_h < 1e-4, worst_h
    trained = tr.end_round()
    d = (trained - hr.mu2_table.detach()).abs()
    assert float(d.max()) <= 2 * 2 * 1e-3 and float((d <= 1e-4 * float(hr.mu2_table.abs().max())).float().mean()) >= 0.999
    assert torch.equal(master.fetch(utts), trained), "owners hold the trained rows"
    if rank == 0:
        print(f"mp_check ok: world {world}, loss rel err {worst:.2e}, sharded-table loss rel err {worst_s:.2e}, "
              f"hierarchical round loss rel err {worst_h:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
