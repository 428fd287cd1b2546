"""GPU: the rows either side of the hot path: posterior extraction over whole utterances
(config 4), hierarchical-sampling table (config 3), checkpoint interop."""
import json
import os

import numpy as np
import pytest
import torch

import pytorch_scalablefhvae_b200 as P
from oracle import fhvae_oracle as O
from util import FP32_RTOL, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model_pair(N=9, mode=P.MODE_F32_SIMT, H=32, F=8, T=6, Z1=8, Z2=16):
    torch.manual_seed(0)
    args = (T * F, [H, H], [H, H], Z1, Z2, [H, H])
    m = P.FHVAE(*args, seg_len=T, num_seqs=N, gemm_mode=mode)
    o = O.FHVAEOracle(*args, seg_len=T, num_seqs=N)
    o.load_state_dict(m.state_dict())
    return m.to(DEV), o


@pytest.mark.parametrize("bs", [7, 64])
def test_extract_posteriors_matches_oracle(bs):
    T, F, shift = 6, 8, 2
    m, o = _model_pair()
    rng = np.random.default_rng(0)
    lengths = [int(v) for v in rng.integers(3, 40, size=11)]          # includes utterances shorter than seg_len
    lengths[3] = 5
    feats = torch.randn(sum(lengths), F, generator=torch.Generator().manual_seed(1))
    out = P.extract_posteriors(m, feats.to(DEV), lengths, seg_shift=shift, batch_size=bs)
    # reference segmenting (datasets.py:176-181) + encoders with eps = 0 + utils.estimate_mu2_dict
    segs, utt = [], []
    off = 0
    for u, l in enumerate(lengths):
        for s in O.segment_starts(l, T, shift):
            segs.append(feats[off + s: off + s + T]); utt.append(u)
        off += l
    x = torch.stack(segs)
    utt = torch.tensor(utt)
    assert torch.equal(out["seg_utt"].cpu(), utt)                       # bit-exact indices
    assert out["nsegs"].tolist() == [max((l - T) // shift + 1, 0) for l in lengths]
    zero = {"z1": torch.zeros(len(x), 8), "z2": torch.zeros(len(x), 16)}
    with torch.no_grad():
        o(x, torch.zeros(len(x), dtype=torch.long), 9, torch.ones(len(x), dtype=torch.long), eps=zero)
    assert_close(out["z2_mu"], o.qz2_x[0], FP32_RTOL, "z2_mu")
    assert_close(out["z1_mu"], o.qz1_x[0], FP32_RTOL, "z1_mu")
    d = O.estimate_mu2_dict([o.qz2_x[0]], [utt])
    for u in range(len(lengths)):
        ref = d[u] if u in d else torch.zeros(16)
        assert_close(out["mu2"][u], ref, FP32_RTOL, f"mu2[{u}]") if u in d else None
        if u not in d:
            assert float(out["mu2"][u].abs().max()) == 0.0


def test_gather_segments_mvn_exact():
    from pytorch_scalablefhvae_b200 import _lib
    from pytorch_scalablefhvae_b200.plan import ptr
    R, F, T = 50, 12, 5
    feats = torch.randn(R, F, device=DEV)
    start = torch.tensor([0, 45, 7, 7], device=DEV)
    mean, std = torch.randn(F, device=DEV), torch.rand(F, device=DEV) + 0.5
    inv = 1.0 / std
    out = torch.empty(4, T, F, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.fn("fhvae_gather_segments")(ptr(feats), ptr(start), None, None, ptr(out), 4, T, F, R, st))
    ref = torch.stack([feats[s:s + T] for s in start.tolist()])
    assert torch.equal(out, ref)
    _lib.check(_lib.fn("fhvae_gather_segments")(ptr(feats), ptr(start), ptr(mean), ptr(inv), ptr(out), 4, T, F, R, st))
    assert_close(out, (ref - mean) * inv, 1e-6, "mvn")


def test_hierarchical_table_roundtrip_and_refresh(golden_dir):
    h = json.load(open(os.path.join(golden_dir, "hier_sample.json")))
    seqlist = [f"utt{i:05d}" for i in range(h["n"])]
    assert P.sample_sequences(seqlist, h["k"], h["seed"]).tolist() == h["sampled"]       # train_model.py:426-428
    N, K, Z = 1000, 50, 16
    table = P.ShardedMu2Table(N, Z, DEV, seed=3)
    master0 = table.shard.clone()
    utts = torch.from_numpy(np.random.RandomState(5).permutation(N)[:K].astype(np.int64))
    cache = table.fetch(utts)
    assert torch.equal(cache, master0[utts.to(DEV)])                                       # exact rows (world 1)
    new = cache + 1.0
    table.write_back(utts, new)
    touched = torch.zeros(N, dtype=torch.bool); touched[utts] = True
    assert torch.equal(table.shard[touched.to(DEV)], (master0 + 1.0)[touched.to(DEV)])
    assert torch.equal(table.shard[~touched.to(DEV)], master0[~touched.to(DEV)])          # untouched rows intact
    # cache refresh with the current encoder == utils.estimate_mu2_dict on the oracle
    m, o = _model_pair(N=K)
    g = torch.Generator().manual_seed(7)
    xs = [torch.randn(20, 6, 8, generator=g) for _ in range(3)]
    labs = [torch.randint(0, K - 5, (20,), generator=g) for _ in range(3)]
    before = m.mu2_table.detach().clone()
    t = table.refresh(m, [x.to(DEV) for x in xs], labs)
    zero = {"z1": torch.zeros(20, 8), "z2": torch.zeros(20, 16)}
    z2s = []
    with torch.no_grad():
        for x in xs:
            o(x, torch.zeros(20, dtype=torch.long), K, torch.ones(20, dtype=torch.long), eps=zero)
            z2s.append(o.qz2_x[0].clone())
    d = O.estimate_mu2_dict(z2s, labs)
    for y in range(K):
        if y in d:
            assert_close(t[y], d[y], FP32_RTOL, f"mu2[{y}]")
        else:
            assert torch.equal(t[y], before[y])                                            # never-seen rows keep their value


def test_checkpoint_interop_with_reference_layout(tmp_path):
    m, o = _model_pair()
    opt = P.FusedAdam(m.parameters(), lr=1e-3, betas=(0.95, 0.999))
    x = torch.randn(5, 6, 8); idx = torch.tensor([0, 8, 3, 3, 1]); ns = torch.tensor([2, 3, 4, 4, 9])
    out = m(x.to(DEV), idx, 9, ns); P.loss_function(out[0], out[1]).backward(); opt.step()
    path = P.save_checkpoint(m, opt, [], {"a": 1}, "run", epoch=3, best_epoch=3, val_lower_bound=-1.0,
                             best_val_lb=-1.0, checkpoint_dir=str(tmp_path))
    assert path.name == "fhvae_run_e3.tar" and (tmp_path / "best_model_fhvae_run_e3.tar").exists()
    ck = torch.load(path, weights_only=False)
    assert set(ck) >= {"best_val_lb", "best_epoch", "epoch", "model_type", "model_params", "optimizer", "state_dict",
                       "summary_vals", "values"}                                            # utils.py:126-145
    assert ck["model_params"] == ([32, 32], [32, 32], 8, 16, [32, 32])
    m2, values, optim_state, start_epoch, best, summ = P.load_checkpoint_file(path, finetune=False)
    assert start_epoch == 5 and values == {"a": 1}                                          # utils.py:89-92 (+2)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a.cpu(), b.cpu()), k
    # optimizer state is torch.optim.Adam-shaped: it loads into the reference's optimizer and back
    ref_opt = torch.optim.Adam(o.parameters(), lr=1e-3, betas=(0.95, 0.999))
    names_m = [n for n, _ in m.named_parameters()]
    names_o = [n for n, _ in o.named_parameters()]
    assert names_m == names_o
    sd = optim_state
    ref_opt.load_state_dict({"state": {i: {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in st.items()}
                                       for i, st in sd["state"].items()}, "param_groups": sd["param_groups"]})
    m2.to(DEV)
    opt2 = P.FusedAdam(m2.parameters(), lr=1e-3, betas=(0.95, 0.999))
    opt2.load_state_dict(optim_state)
    assert opt2.steps_taken() == 1
    a, b = opt._flat_state[id(m)], opt2._flat_state[id(m2)]
    assert torch.equal(a["m"], b["m"]) and torch.equal(a["v"], b["v"])
    # a REFERENCE SimpleFHVAE checkpoint (no table, no input_size) loads too
    sref = O.SimpleFHVAEOracle(24, [8, 8], [8, 8], 8, 8, [8, 8], num_seqs=4)
    ref_sd = {k: v for k, v in sref.state_dict().items() if k != "mu2_table"}
    torch.save({"model_type": "simple_fhvae", "model_params": ([8, 8], [8, 8], 8, 8, [8, 8]), "state_dict": ref_sd,
                "optimizer": {}, "epoch": 0, "best_val_lb": 0, "summary_vals": [], "values": {}, "best_epoch": 0},
               tmp_path / "ref.tar")
    m3 = P.load_checkpoint_file(tmp_path / "ref.tar", finetune=True, input_size=24, num_seqs=4)[0]
    assert torch.equal(m3.state_dict()["pre_decoder.fc2.linear.weight"], ref_sd["pre_decoder.fc2.linear.weight"])


def test_hierarchical_round_end_to_end():
    """BASELINE config 3, one rank: sample (bit-exact) -> fetch from the master table -> K-row ACTIVE table sharded
    inside the train step -> two steps == the oracle training on the fetched rows -> write-back touches exactly the
    sampled master rows."""
    Nm, K, Z2, B, T, F = 300, 24, 16, 10, 6, 8
    m, o = _model_pair(N=K)
    opt = P.FusedAdam(m.parameters(), lr=1e-3, betas=(0.95, 0.999))
    master = P.ShardedMu2Table(Nm, Z2, DEV, seed=3)
    master0 = master.shard.clone()
    tr = P.HierarchicalTrainer(m, opt, master, K)
    utts = tr.begin_round(seed=11)
    assert utts.tolist() == np.random.RandomState(11).choice(np.arange(Nm), K, replace=False).tolist()
    assert torch.equal(m.mu2_table.detach(), master0[utts.to(DEV)])                    # exact rows, label = position
    with torch.no_grad():
        o.mu2_table.copy_(master0[utts.to(DEV)].cpu())
    oopt = O.make_adam(o.parameters())
    g = torch.Generator().manual_seed(4)
    for step in range(2):
        x = torch.randn(B, T, F, generator=g)
        lab = torch.randint(0, K, (B,), generator=g)
        ns = torch.randint(1, 50, (B,), generator=g)
        eps = {"z1": torch.randn(B, 8, generator=g), "z2": torch.randn(B, 16, generator=g)}
        loss = tr.train_step(x.to(DEV), lab.to(DEV), ns.to(DEV), 10.0, eps=eps)
        rl, _ = O.train_step(o, oopt, x, lab, K, ns, 10.0, eps=eps)
        assert_close(loss, rl, FP32_RTOL, f"loss step {step}")
    cache = tr.end_round()
    assert_close(cache, o.mu2_table.detach(), FP32_RTOL, "trained rows")
    touched = torch.zeros(Nm, dtype=torch.bool, device=DEV)
    touched[utts.to(DEV)] = True
    assert torch.equal(master.shard[utts.to(DEV)], cache)                              # owners hold the trained rows
    assert torch.equal(master.shard[~touched], master0[~touched])                      # all other rows bit-identical
    # next round: fresh Adam moments for the (new) active rows, MAP refresh available
    xs = [torch.randn(B, T, F, generator=g).to(DEV) for _ in range(2)]
    labs = [torch.randint(0, K, (B,), generator=g) for _ in range(2)]
    utts2 = tr.begin_round(seed=12, refresh=(xs, labs))
    st = opt._flat_state[id(m)]
    o_t = m._off["mu2_table"]
    assert float(st["m"][o_t:].abs().max()) == 0.0 and float(st["v"][o_t:].abs().max()) == 0.0
    assert utts2.tolist() != utts.tolist()


def test_extract_posteriors_sharded_covers_all_utterances():
    T, F, shift = 6, 8, 2
    m, _ = _model_pair()
    rng = np.random.default_rng(1)
    lengths = [int(v) for v in rng.integers(6, 40, size=13)]
    feats = torch.randn(sum(lengths), F, generator=torch.Generator().manual_seed(2)).to(DEV)
    offs = np.concatenate([[0], np.cumsum(lengths)])
    feats_of = lambda ids: torch.cat([feats[offs[u]:offs[u + 1]] for u in ids])
    whole = {k: v.clone() for k, v in P.extract_posteriors(m, feats, lengths, seg_shift=shift, batch_size=16).items()}
    seen, mu2 = [], torch.zeros_like(whole["mu2"])
    for rank in range(3):
        out = P.extract_posteriors_sharded(m, feats_of, lengths, rank, 3, seg_shift=shift, batch_size=16)
        seen += out["utts"].tolist()
        mu2[out["utts"].to(DEV)] = out["mu2"]
    assert sorted(seen) == list(range(13))
    assert_close(mu2, whole["mu2"], 1e-5, "mu2")             # per-utterance results do not depend on the sharding
                                                             # (batch boundaries differ: summation order only)


def test_default_extraction_batch_is_a_whole_number_of_recurrence_launches():
    """Forward-only batches are sized so that no recurrence launch runs with a partial set of groups
    (fhvae_lstm_wave_rows_per_launch: groups x 32 rows that fit the device), and ragged tails run padded."""
    rows = P._lib.fn("fhvae_lstm_wave_rows_per_launch")(256, 2)
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    assert rows == min(sm // 16, 32) * 32 and rows > 0          # 8 CTAs per group x 2 layers, one CTA per SM
    assert P._lib.fn("fhvae_lstm_wave_rows_per_launch")(200, 2) == 0
    m = P.FHVAE(1600, [256, 256], [256, 256], 32, 32, [256, 256], seg_len=20, num_seqs=10, gemm_mode=P.MODE_BF16X3)
    bs = P.inference.default_batch_size(m)
    assert bs % rows == 0 and 2048 - rows < bs <= 2048
    small = P.SimpleFHVAE(1600, [128, 128], [128, 128], 16, 16, [128, 128], num_seqs=10)
    assert P.inference.default_batch_size(small) == 2048
    # a ragged extraction batch of the LSTM model (51 segments) is served by the padded plan
    m = m.to(DEV)
    enc = m.encode(torch.randn(51, 20, 80, device=DEV))
    assert enc["z2_mu"].shape == (51, 32) and list(m._plans.values())[0].B == 64
