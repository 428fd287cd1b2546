"""CPU: pin the oracle (oracle/fhvae_oracle.py) against fixtures produced by the UNMODIFIED
reference (oracle/make_golden.py).  SURVEY.md §8c."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import fhvae_oracle as O

RTOL = 2e-6   # fp32-vs-fp32 on CPU, same op order up to summation order


def _load_tiny(golden_dir):
    d = np.load(os.path.join(golden_dir, "simple_fhvae_tiny.npz"))
    return {k: (torch.from_numpy(d[k]) if k != "meta" else d[k]) for k in d.files}


def _tiny_model(g, **kw):
    T, F, B, N, Z, H = [int(v) for v in g["meta"]]
    m = O.SimpleFHVAEOracle(T * F, [H, H], [H, H], Z, Z, [H, H], num_seqs=N, **kw)
    sd = {k[2:]: v for k, v in g.items() if k.startswith("w:")}
    sd["mu2_table"] = g["table"]
    m.load_state_dict(sd, strict=True)
    return m


def _eps(g):
    return {"z2": g["eps_z2"], "z1": g["eps_z1"], "x": g["eps_x"]}


@pytest.mark.parametrize("mode", ["ref_compat", "debugged"])
def test_simple_forward_matches_reference(golden_dir, mode):
    g = _load_tiny(golden_dir)
    kw = dict(detach_px=True, prior_grad=False, ref_log_qy=True) if mode == "ref_compat" else {}
    m = _tiny_model(g, **kw)
    out = m(g["x"], g["idx"], m.mu2_table.shape[0], g["nsegs"], eps=_eps(g))
    names = ["lower_bound", "log_qy", "log_px_z", "neg_kld_z1", "neg_kld_z2", "log_pmu2"]
    for n, o in zip(names, out):
        ref = g["out_" + n]
        if n == "log_qy" and mode == "debugged":
            o = -o.mean()                       # reference returns mean(+CE), simple_fhvae.py:122
        torch.testing.assert_close(o.detach(), ref, rtol=RTOL, atol=1e-6, msg=n)


def test_simple_gradients_match_reference_as_is(golden_dir):
    """ref_compat mode reproduces the reference's detach placement -> its gradients exactly."""
    g = _load_tiny(golden_dir)
    m = _tiny_model(g, detach_px=True, prior_grad=False, ref_log_qy=True)
    out = m(g["x"], g["idx"], m.mu2_table.shape[0], g["nsegs"], eps=_eps(g))
    loss = O.loss_function(out[0], out[1], 10.0)
    torch.testing.assert_close(loss.detach(), g["loss"], rtol=RTOL, atol=1e-6)
    loss.backward()
    n_checked = 0
    for k, p in m.named_parameters():
        if k == "mu2_table":
            torch.testing.assert_close(p.grad, g["grad_table"], rtol=1e-5, atol=1e-7)
            continue
        if "g:" + k in g:
            torch.testing.assert_close(p.grad, g["g:" + k], rtol=1e-5, atol=1e-7, msg=k)
            n_checked += 1
        else:                                   # decoder: no gradient in the reference (F6)
            assert k.startswith(("pre_decoder", "dec_gauss_layer"))
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
    assert n_checked == 16


def test_config0_kat_regenerated(golden_dir):
    """SURVEY.md §4 known-answer vector, reproduced by the oracle with the reference's RNG draw order."""
    kat = json.load(open(os.path.join(golden_dir, "simple_fhvae_c0_kat.json")))
    if kat["torch"] != torch.__version__:
        pytest.skip("RNG streams / default init pinned to torch " + kat["torch"])
    torch.manual_seed(0)
    m = O.SimpleFHVAEOracle(1600, num_seqs=1000, detach_px=True, prior_grad=False, ref_log_qy=True)
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(64, 20, 80, generator=gen)
    idx = torch.randint(0, 1000, (64,), generator=gen)
    nsegs = torch.randint(1, 200, (64,), generator=gen)
    assert idx[:8].tolist() == kat["idx_head"]
    torch.manual_seed(2)                         # simple_fhvae.py:51 then :215 x3 (z2, z1, x)
    table = torch.empty(1000, 16).normal_(mean=0, std=1.0)
    eps = {"z2": torch.randn(64, 16), "z1": torch.randn(64, 16), "x": torch.randn(64, 1600)}
    with torch.no_grad():
        m.mu2_table.copy_(table)
    out = m(x, idx, 1000, nsegs, eps=eps)
    loss = O.loss_function(out[0], out[1], 10.0)
    loss.backward()
    got = {"mean_lower_bound": out[0].mean(), "log_qy": out[1], "mean_log_px_z": out[2].mean(),
           "mean_neg_kld_z1": out[3].mean(), "mean_neg_kld_z2": out[4].mean(),
           "mean_log_pmu2": out[5].mean(), "loss_alpha10": loss,
           "gnorm_z2_pre_encoder_fc1_w": m.z2_pre_encoder.fc1.linear.weight.grad.norm(),
           "gnorm_z1_gauss_mulayer_w": m.z1_gauss_layer.mulayer.weight.grad.norm()}
    for k, v in got.items():
        assert float(v) == pytest.approx(kat[k], rel=2e-5), k


def test_hierarchical_sample_bit_exact(golden_dir):
    h = json.load(open(os.path.join(golden_dir, "hier_sample.json")))
    seqlist = [f"utt{i:05d}" for i in range(h["n"])]
    assert O.hierarchical_sample(seqlist, h["k"], h["seed"]).tolist() == h["sampled"]
    # permutation identity of SURVEY.md Appendix D
    perm = np.random.RandomState(h["seed"]).permutation(h["n"])[: h["k"]]
    assert np.asarray(seqlist)[perm].tolist() == h["sampled"]


def test_estimate_mu2_batched_equals_dict_loop():
    g = torch.Generator().manual_seed(3)
    z2 = torch.randn(40, 8, generator=g)
    idx = torch.randint(0, 7, (40,), generator=g)
    d = O.estimate_mu2_dict([z2[:13], z2[13:]], [idx[:13], idx[13:]])
    t, n = O.estimate_mu2_table(z2, idx, 7)
    for y, v in d.items():
        torch.testing.assert_close(t[y].float(), v, rtol=1e-6, atol=1e-7)
        assert int(n[y]) == int((idx == y).sum())


def test_fhvae_oracle_runs_and_shapes():
    torch.manual_seed(0)
    m = O.FHVAEOracle(4 * 6, [16, 16], [16, 16], 8, 8, [16, 16], seg_len=4, num_seqs=9)
    x = torch.randn(5, 4, 6)
    idx = torch.tensor([0, 8, 3, 3, 1])
    out = m(x, idx, 9, torch.tensor([2, 3, 4, 4, 9]))
    assert [tuple(o.shape) for o in out] == [(5,)] * 6
    O.loss_function(out[0], out[1]).backward()
    assert all(p.grad is not None for p in m.parameters())
    keys = set(m.state_dict())
    assert "z2_pre_encoder.lstm.weight_hh_l1" in keys and "dec_gauss_layer.logvar_layer.bias" in keys


def test_segment_arithmetic():
    assert O.segment_starts(200).tolist() == list(range(0, 184, 8))   # (200-20)//8+1 = 23 segments
    assert len(O.segment_starts(20)) == 1 and len(O.segment_starts(19)) == 0


# ---------------------------------------------------------------------------------------------------------------
# O3 (FHVAE, LSTM): the reference's fhvae.py:4-14 is a stub, so O3 is "parity unpinned" by the reference.  It is
# pinned twice by us instead: (1) against a hand-written fp64 numpy LSTM cell + closed-form loss / BPTT (SURVEY.md
# Appendix C) that shares no code with torch.nn.LSTM or autograd, (2) against a committed golden of itself.
# ---------------------------------------------------------------------------------------------------------------
def _o3_small(seed=21):
    T, F, B, N, Z1, Z2, H = 6, 8, 10, 13, 8, 16, 32
    torch.manual_seed(seed)
    m = O.FHVAEOracle(T * F, [H, H], [H, H], Z1, Z2, [H, H], seg_len=T, num_seqs=N).double()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, T, F, generator=g)
    idx = torch.randint(0, N, (B,), generator=g)
    nsegs = torch.randint(1, 120, (B,), generator=g)
    eps = {"z2": torch.randn(B, Z2, generator=g), "z1": torch.randn(B, Z1, generator=g)}
    return m, x, idx, nsegs, eps, N


def test_o3_forward_matches_handwritten_fp64_cell():
    m, x, idx, nsegs, eps, N = _o3_small()
    out = m(x.double(), idx, N, nsegs, eps=eps)
    loss = O.loss_function(out[0], out[1], 10.0)
    ref = O.fhvae_forward_fp64(m, x, idx, nsegs, eps)
    names = ["lower_bound", "log_qy", "log_px_z", "neg_kld_z1", "neg_kld_z2", "log_pmu2"]
    for n, o in zip(names, out):
        np.testing.assert_allclose(o.detach().numpy(), ref[n], rtol=1e-10, atol=1e-10, err_msg=n)
    assert float(loss) == pytest.approx(float(ref["loss"]), rel=1e-12)


@pytest.mark.parametrize("layers", [1, 2, 3])
def test_nn_lstm_matches_handwritten_cell_and_bptt(layers):
    """nn.LSTM forward + autograd == hand-written cell + closed-form BPTT (outputs, final h of every layer, dx, every
    weight / bias gradient), fp64, 1e-10."""
    B, T, In, H = 5, 7, 6, 12
    torch.manual_seed(3)
    lstm = torch.nn.LSTM(In, H, num_layers=layers, batch_first=True).double()
    x = torch.randn(B, T, In, dtype=torch.float64, requires_grad=True)
    out, (hn, _) = lstm(x)
    g = torch.Generator().manual_seed(4)
    d_out = torch.randn(B, T, H, generator=g, dtype=torch.float64)
    d_fin = torch.randn(layers, B, H, generator=g, dtype=torch.float64)
    ((out * d_out).sum() + (hn * d_fin).sum()).backward()
    w = [tuple(getattr(lstm, f"{n}_l{l}").detach().numpy() for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"))
         for l in range(layers)]
    o2, fin, cache = O.lstm_stack_fp64(x.detach().numpy(), w)
    np.testing.assert_allclose(out.detach().numpy(), o2, rtol=1e-10, atol=1e-12)
    for l in range(layers):
        np.testing.assert_allclose(hn[l].detach().numpy(), fin[l], rtol=1e-10, atol=1e-12)
    dx, grads = O.lstm_stack_bwd_fp64(d_out.numpy(), [d_fin[l].numpy() for l in range(layers)], w, cache)
    np.testing.assert_allclose(x.grad.numpy(), dx, rtol=1e-9, atol=1e-12)
    for l in range(layers):
        for n, gr in zip(("weight_ih", "weight_hh", "bias_ih", "bias_hh"), grads[l]):
            np.testing.assert_allclose(getattr(lstm, f"{n}_l{l}").grad.numpy(), gr, rtol=1e-9, atol=1e-12,
                                       err_msg=f"{n}_l{l}")


def test_o3_reproduces_its_golden(golden_dir):
    """tests/golden/fhvae_o3_small.npz (oracle/make_golden.py): the oracle on THIS box gives the committed values."""
    d = np.load(os.path.join(golden_dir, "fhvae_o3_small.npz"))
    g = {k: torch.from_numpy(d[k]) for k in d.files if k != "meta"}
    T, F, B, N, Z1, Z2, H = [int(v) for v in d["meta"]]
    m = O.FHVAEOracle(T * F, [H, H], [H, H], Z1, Z2, [H, H], seg_len=T, num_seqs=N)
    m.load_state_dict({k[2:]: v for k, v in g.items() if k.startswith("w:")}, strict=True)
    eps = {"z1": g["eps_z1"], "z2": g["eps_z2"]}
    out = m(g["x"], g["idx"], N, g["nsegs"], eps=eps)
    names = ["lower_bound", "log_qy", "log_px_z", "neg_kld_z1", "neg_kld_z2", "log_pmu2"]
    for n, o in zip(names, out):
        torch.testing.assert_close(o.detach(), g["out_" + n], rtol=2e-5, atol=1e-5, msg=n)
    loss = O.loss_function(out[0], out[1], 10.0)
    loss.backward()
    torch.testing.assert_close(loss.detach(), g["loss"], rtol=2e-5, atol=1e-5)
    for k, p in m.named_parameters():
        e = float((p.grad - g["g:" + k]).abs().max()) / (float(g["g:" + k].abs().max()) + 1e-30)
        assert e < 2e-5, (k, e)
    # and the golden itself agrees with the hand-written fp64 cell (fp32 oracle vs fp64 numpy: 1e-5)
    ref = O.fhvae_forward_fp64(m, g["x"], g["idx"], g["nsegs"], eps)
    for n in names:
        np.testing.assert_allclose(g["out_" + n].numpy(), ref[n], rtol=2e-5, atol=2e-5, err_msg=n)
