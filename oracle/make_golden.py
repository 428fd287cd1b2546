"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden.py

It imports ``/root/reference/simple_fhvae.py`` as-is, and -- following SURVEY.md Appendix D --
injects the mu2 table (class-level patch of ``SimpleFHVAE.mu2_lookup``, simple_fhvae.py:39-54)
and the eps draws (class-level patch of ``GaussianLayer.forward``, simple_fhvae.py:211-216, consumed
in call order z2, z1, x -- simple_fhvae.py:91,95,99), so that ``forward`` is deterministic given
the stored tensors.  No reference source is copied; only its outputs are stored.

Fixtures written:
* ``simple_fhvae_tiny.npz``  -- every weight / input / eps / table tensor, the six forward outputs
  and the as-is gradients (encoders + table; the reference's decoder gets none) of a tiny config.
* ``simple_fhvae_c0_kat.json`` -- the seeded config-0 known-answer scalars of SURVEY.md §4,
  regenerated with the installed torch (uninjected: the reference's own RNG draws).
* ``fhvae_o3_small.npz`` -- O3 (the nn.LSTM restatement; the reference's fhvae.py is a stub, so this one is NOT
  produced by the reference): weights, inputs, eps, six outputs, loss and every gradient of a small 2x32 LSTM
  FHVAE.  It makes the authoring box and the GPU box provably use the same oracle values (a torch upgrade that moved
  nn.LSTM would fail tests/test_oracle.py::test_o3_reproduces_its_golden).  Does not need /root/reference.
* ``hier_sample.json`` -- np.random.choice(seqlist, K, replace=False) under np.random.seed(s)
  (train_model.py:426-428) for a 1000-utterance list, K=50.
"""
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    sys.path.insert(0, REF)
    import simple_fhvae as ref  # noqa
    return ref


class _Inject:
    """Context manager that patches table + eps injection into the reference classes."""

    def __init__(self, ref, table, eps_list):
        self.ref, self.table, self.eps = ref, table, list(eps_list)

    def __enter__(self):
        ref, table, eps = self.ref, self.table, self.eps
        self._lookup = ref.SimpleFHVAE.mu2_lookup
        self._gfwd = ref.GaussianLayer.forward

        def lookup(self_, mu_idx, z2_dim, num_seqs, init_std=1.0):
            return table, table[mu_idx]

        def gfwd(self_, h):
            mu = self_.mulayer(h)
            logvar = self_.logvar_layer(h)
            e = eps.pop(0)
            return mu, logvar, mu + e.reshape(mu.shape) * torch.exp(0.5 * logvar)

        ref.SimpleFHVAE.mu2_lookup = lookup
        ref.GaussianLayer.forward = gfwd
        return self

    def __exit__(self, *a):
        self.ref.SimpleFHVAE.mu2_lookup = self._lookup
        self.ref.GaussianLayer.forward = self._gfwd


def tiny(ref):
    T, F, B, N, Z, H = 4, 6, 5, 12, 16, 32           # Z must be 16: simple_fhvae.py:53 hard-codes it
    torch.manual_seed(11)
    m = ref.SimpleFHVAE(T * F, [H, H], [H, H], Z, Z, [H, H])
    g = torch.Generator().manual_seed(12)
    x = torch.randn(B, T, F, generator=g)
    idx = torch.tensor([3, 0, 11, 3, 7])              # duplicate + first + last row
    nsegs = torch.tensor([5, 1, 17, 5, 9])
    table = torch.randn(N, Z, generator=g).requires_grad_(True)
    eps = [torch.randn(B, Z, generator=g), torch.randn(B, Z, generator=g), torch.randn(B, T * F, generator=g)]
    with _Inject(ref, table, eps):
        out = m(x, idx, N, nsegs)
    lb, log_qy = out[0], out[1]
    loss = -1 * torch.mean(lb + 10.0 * log_qy)        # train_model.py:251, alpha 10
    loss.backward()
    d = {"x": x, "idx": idx, "nsegs": nsegs, "table": table.detach(),
         "eps_z2": eps[0], "eps_z1": eps[1], "eps_x": eps[2],
         "out_lower_bound": out[0], "out_log_qy": out[1], "out_log_px_z": out[2],
         "out_neg_kld_z1": out[3], "out_neg_kld_z2": out[4], "out_log_pmu2": out[5],
         "loss": loss, "grad_table": table.grad}
    for k, v in m.state_dict().items():
        d["w:" + k] = v
    for k, p in m.named_parameters():
        if p.grad is not None:
            d["g:" + k] = p.grad
    d["meta"] = np.array([T, F, B, N, Z, H])
    np.savez_compressed(os.path.join(OUT, "simple_fhvae_tiny.npz"),
                        **{k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in d.items()})
    print("tiny: loss", float(loss), "grads for", sum(k.startswith("g:") for k in d), "tensors")


def config0_kat(ref):
    torch.manual_seed(0)
    m = ref.SimpleFHVAE(1600)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(64, 20, 80, generator=g)
    idx = torch.randint(0, 1000, (64,), generator=g)
    nsegs = torch.randint(1, 200, (64,), generator=g)
    torch.manual_seed(2)
    out = m(x, idx, 1000, nsegs)
    loss = -1 * torch.mean(out[0] + 10.0 * out[1])
    loss.backward()
    kat = {
        "torch": torch.__version__,
        "mean_lower_bound": float(out[0].mean()), "log_qy": float(out[1]),
        "mean_log_px_z": float(out[2].mean()), "mean_neg_kld_z1": float(out[3].mean()),
        "mean_neg_kld_z2": float(out[4].mean()), "mean_log_pmu2": float(out[5].mean()),
        "loss_alpha10": float(loss),
        "gnorm_z2_pre_encoder_fc1_w": float(m.z2_pre_encoder.fc1.linear.weight.grad.norm()),
        "gnorm_z1_gauss_mulayer_w": float(m.z1_gauss_layer.mulayer.weight.grad.norm()),
        "idx_head": idx[:8].tolist(),
        "decoder_grads_none": all(p.grad is None for n, p in m.named_parameters()
                                  if n.startswith(("pre_decoder", "dec_gauss_layer"))),
    }
    with open(os.path.join(OUT, "simple_fhvae_c0_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("c0 KAT:", kat)


def hier():
    seqlist = [f"utt{i:05d}" for i in range(1000)]
    np.random.seed(5)
    s = np.random.choice(seqlist, 50, replace=False)   # train_model.py:426-428
    with open(os.path.join(OUT, "hier_sample.json"), "w") as f:
        json.dump({"seed": 5, "n": 1000, "k": 50, "sampled": s.tolist()}, f)


def fhvae_o3_small():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import fhvae_oracle as O
    T, F, B, N, Z1, Z2, H = 6, 8, 10, 13, 8, 16, 32
    torch.manual_seed(21)
    m = O.FHVAEOracle(T * F, [H, H], [H, H], Z1, Z2, [H, H], seg_len=T, num_seqs=N)
    g = torch.Generator().manual_seed(22)
    x = torch.randn(B, T, F, generator=g)
    idx = torch.tensor([3, 0, 12, 3, 7, 7, 7, 1, 12, 5])      # duplicates + first + last row
    nsegs = torch.randint(1, 120, (B,), generator=g)
    eps = {"z2": torch.randn(B, Z2, generator=g), "z1": torch.randn(B, Z1, generator=g)}
    out = m(x, idx, N, nsegs, eps=eps)
    loss = O.loss_function(out[0], out[1], 10.0)
    loss.backward()
    d = {"x": x, "idx": idx, "nsegs": nsegs, "eps_z2": eps["z2"], "eps_z1": eps["z1"], "loss": loss}
    for n, o in zip(["lower_bound", "log_qy", "log_px_z", "neg_kld_z1", "neg_kld_z2", "log_pmu2"], out):
        d["out_" + n] = o
    for k, v in m.state_dict().items():
        d["w:" + k] = v
    for k, p in m.named_parameters():
        d["g:" + k] = p.grad
    d["meta"] = np.array([T, F, B, N, Z1, Z2, H])
    np.savez_compressed(os.path.join(OUT, "fhvae_o3_small.npz"),
                        **{k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in d.items()})
    print("fhvae_o3_small: loss", float(loss), "torch", torch.__version__)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    fhvae_o3_small()
    if os.path.isdir(REF):
        ref = _import_reference()
        tiny(ref)
        config0_kat(ref)
        hier()
