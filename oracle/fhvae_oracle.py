"""CPU oracle for the ScalableFHVAE train / inference step.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
CUDA package (``pytorch_scalablefhvae_b200``) must never import anything from ``oracle/``.

What it restates (file:line into the reference, BurnhamG/PyTorch-ScalableFHVAE):

* ``log_gauss`` / ``kld``              -- simple_fhvae.py:56-69
* ELBO assembly                        -- simple_fhvae.py:106-116
* discriminative logits / CE           -- simple_fhvae.py:119-122
* ``loss_function``                    -- train_model.py:243-251
* ``SimpleFHVAEOracle`` (O2)           -- simple_fhvae.py:8-124 + sub-modules :127-244
* ``FHVAEOracle`` (O3, nn.LSTM)        -- fhvae.py:4-14 is a stub (raises); structure follows
                                          simple_fhvae.py:86-124 with recurrent blocks per the
                                          docstrings simple_fhvae.py:138-149,168-177,220-231
* ``estimate_mu2_dict`` / batched form -- utils.py:45-60
* hierarchical utterance sampling      -- train_model.py:424-428

Parity pinning
--------------
* SimpleFHVAE (O2): PINNED.  ``oracle/make_golden.py`` imports the unmodified reference
  ``simple_fhvae.py`` in the authoring container, injects the table and the eps draws, and
  commits inputs + outputs + as-is gradients under ``tests/golden/``;
  ``tests/test_oracle.py`` checks O2 against them (forward values in every mode, gradients in
  ``ref_compat`` mode which reproduces the reference's detach placement).
* FHVAE (O3): **parity unpinned** by the reference -- ``fhvae.py`` implements nothing, the
  reference holds no tests/golden vectors.  O3 shares *the same loss functions* as O2 (pinned)
  and uses ``torch.nn.LSTM`` for the recurrent blocks, so the only unpinned part is the
  architecture choice documented in DESIGN.md.

Deliberate differences from the as-is reference, each switchable:
* the mu2 table is a persistent ``nn.Parameter`` instead of fresh normals every forward (A1);
* ``detach_px=False`` lets the decoder train (A2);  ``detach_px=True`` reproduces the reference;
* ``prior_grad=True`` lets log p(mu2) reach the table (A3);  ``False`` reproduces the reference;
* ``log_qy`` is returned per segment as ``-CE_b`` (A4); ``ref_log_qy=True`` returns the
  reference's positive mean scalar.
"""
from __future__ import annotations

import math
from collections import defaultdict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

LOG_2PI = math.log(2.0 * math.pi)
PZ1 = (0.0, math.log(1.0 ** 2))      # simple_fhvae.py:22
PMU2 = (0.0, math.log(1.0 ** 2))     # simple_fhvae.py:23
PZ2_LOGVAR = math.log(0.5 ** 2)      # simple_fhvae.py:88


# --------------------------------------------------------------------------------------
# loss maths
# --------------------------------------------------------------------------------------
def log_gauss(x, mu=0.0, logvar=0.0):
    """log N(x; mu, exp(logvar)) -- simple_fhvae.py:56-60."""
    if not torch.is_tensor(logvar):
        inv = math.exp(-logvar)
        return -0.5 * (LOG_2PI + logvar + (x - mu) ** 2 * inv)
    return -0.5 * (LOG_2PI + logvar + (x - mu) ** 2 / torch.exp(logvar))


def kld(p_mu, p_logvar, q_mu, q_logvar: float):
    """KL(p || q) for diagonal Gaussians, q_logvar a python scalar -- simple_fhvae.py:62-69."""
    return -0.5 * (
        1 + p_logvar - q_logvar - ((p_mu - q_mu) ** 2 + torch.exp(p_logvar)) / math.exp(q_logvar)
    )


def elbo_terms(x, x_mu, x_logvar, z1_mu, z1_logvar, z2_mu, z2_logvar, mu2, num_segs,
               detach_px: bool = False, prior_grad: bool = True):
    """simple_fhvae.py:106-116 -> (lower_bound, log_px_z, neg_kld_z1, neg_kld_z2, log_pmu2), each (B,)."""
    mu2_p = mu2 if prior_grad else mu2.detach()
    log_pmu2 = torch.sum(log_gauss(mu2_p, PMU2[0], PMU2[1]), dim=1)
    neg_kld_z2 = -1 * torch.sum(kld(z2_mu, z2_logvar, mu2, PZ2_LOGVAR), dim=1)
    neg_kld_z1 = -1 * torch.sum(kld(z1_mu, z1_logvar, PZ1[0], PZ1[1]), dim=1)
    if detach_px:
        x_mu, x_logvar = x_mu.detach(), x_logvar.detach()
    log_px_z = torch.sum(log_gauss(x, x_mu, x_logvar), dim=(1, 2))
    if not torch.is_tensor(num_segs):
        num_segs = torch.as_tensor(num_segs)
    lower_bound = log_px_z + neg_kld_z1 + neg_kld_z2 + log_pmu2 / num_segs.to(log_pmu2.device)
    return lower_bound, log_px_z, neg_kld_z1, neg_kld_z2, log_pmu2


def disc_logits(z2_mu, mu2_table):
    """simple_fhvae.py:119-121 -- materialises (B, N, Z); fine at oracle sizes."""
    d = z2_mu.unsqueeze(1) - mu2_table.unsqueeze(0)
    return torch.sum(-1 * d ** 2 / (2 * math.exp(PZ2_LOGVAR)), dim=-1)


def log_qy_per_segment(z2_mu, mu2_table, mu_idx):
    """Per-segment log q(i | z2) = -CE_b  (the reference returns mean(+CE), simple_fhvae.py:122)."""
    logits = disc_logits(z2_mu, mu2_table)
    return -torch.nn.functional.cross_entropy(logits, mu_idx, reduction="none")


def loss_function(lower_bound, log_qy, alpha: float = 10.0):
    """train_model.py:243-251."""
    return -1 * torch.mean(lower_bound + alpha * log_qy)


def reparam(mu, logvar, eps):
    """simple_fhvae.py:213-216 with the eps draw injected."""
    return mu + eps * torch.exp(0.5 * logvar)


# --------------------------------------------------------------------------------------
# SimpleFHVAE restatement (O2).  Sub-module / parameter names match the reference so a
# reference state_dict loads 1:1 (strict=False only because of the extra mu2_table).
# --------------------------------------------------------------------------------------
class _FC(nn.Module):                       # VariableLinearLayer, simple_fhvae.py:127-134
    def __init__(self, i, o):
        super().__init__()
        self.linear = nn.Linear(i, o)

    def forward(self, x):
        return torch.relu(self.linear(x))


class _Pre(nn.Module):                      # Latent{Seg,Seq}PreEncoder / PreDecoder :137-190,:219-244
    def __init__(self, i, hus):
        super().__init__()
        self.fc1 = _FC(i, hus[0])
        self.fc2 = _FC(hus[0], hus[1])

    def forward(self, x):
        return self.fc2(self.fc1(x))


class _Gauss(nn.Module):                    # GaussianLayer :193-216 (eps injected)
    def __init__(self, i, d):
        super().__init__()
        self.mulayer = nn.Linear(i, d)
        self.logvar_layer = nn.Linear(i, d)

    def forward(self, h):
        return self.mulayer(h), self.logvar_layer(h)


def _draw_eps(eps: Optional[Dict[str, torch.Tensor]], key: str, like: torch.Tensor):
    if eps is not None and key in eps:
        return eps[key].to(like.dtype).reshape(like.shape)
    return torch.randn_like(like)


class _FHVAEBase(nn.Module):
    """Shared forward tail: ELBO + discriminative term, attribute surface of SURVEY.md §8b."""

    def _finish(self, x, mu_idx, num_segs, z1, z2, px, mu2):
        (z1_mu, z1_lv, z1_s), (z2_mu, z2_lv, z2_s), (x_mu, x_lv) = z1, z2, px
        lb, log_px_z, nk1, nk2, log_pmu2 = elbo_terms(
            x, x_mu, x_lv, z1_mu, z1_lv, z2_mu, z2_lv, mu2, num_segs,
            detach_px=self.detach_px, prior_grad=self.prior_grad)
        log_qy = log_qy_per_segment(z2_mu, self.mu2_table, mu_idx)
        if self.ref_log_qy:
            log_qy = -log_qy.mean()
        self.qz1_x, self.qz2_x, self.px_z = [z1_mu, z1_lv], [z2_mu, z2_lv], [x_mu, x_lv]
        self.pz2 = [mu2, np.float32(PZ2_LOGVAR)]
        self.z1_sample, self.z2_sample = z1_s, z2_s
        return lb, log_qy, log_px_z, nk1, nk2, log_pmu2


class SimpleFHVAEOracle(_FHVAEBase):
    def __init__(self, input_size, z1_hus=(128, 128), z2_hus=(128, 128), z1_dim=16, z2_dim=16,
                 x_hus=(128, 128), num_seqs=1000, detach_px=False, prior_grad=True, ref_log_qy=False):
        super().__init__()
        self.model = "simple_fhvae"
        self.pz1 = [0.0, np.float32(PZ1[1])]
        self.pmu2 = [0.0, np.float32(PMU2[1])]
        self.z1_hus, self.z2_hus, self.x_hus = list(z1_hus), list(z2_hus), list(x_hus)
        self.z1_dim, self.z2_dim = z1_dim, z2_dim
        self.detach_px, self.prior_grad, self.ref_log_qy = detach_px, prior_grad, ref_log_qy
        # simple_fhvae.py:31 sizes this with z1_dim; the concatenated tensor is z2 (Appendix A6)
        self.z1_pre_encoder = _Pre(input_size + z2_dim, self.z1_hus)
        self.z2_pre_encoder = _Pre(input_size, self.z2_hus)
        self.z1_gauss_layer = _Gauss(self.z1_hus[1], z1_dim)
        self.z2_gauss_layer = _Gauss(self.z2_hus[1], z2_dim)
        self.pre_decoder = _Pre(z1_dim + z2_dim, self.x_hus)
        self.dec_gauss_layer = _Gauss(self.x_hus[1], input_size)
        self.mu2_table = nn.Parameter(torch.randn(num_seqs, z2_dim))   # :51, init_std=1.0

    def forward(self, x, mu_idx, num_seqs, num_segs, eps=None):
        B = x.shape[0]
        mu2 = self.mu2_table[mu_idx]                                    # :53
        xf = x.reshape(B, -1)
        z2_mu, z2_lv = self.z2_gauss_layer(self.z2_pre_encoder(xf))     # :90-91
        z2_s = reparam(z2_mu, z2_lv, _draw_eps(eps, "z2", z2_mu))
        z1_mu, z1_lv = self.z1_gauss_layer(self.z1_pre_encoder(torch.cat([xf, z2_s], -1)))  # :94-95
        z1_s = reparam(z1_mu, z1_lv, _draw_eps(eps, "z1", z1_mu))
        x_mu, x_lv = self.dec_gauss_layer(self.pre_decoder(torch.cat([z1_s, z2_s], -1)))    # :98-99
        x_mu, x_lv = x_mu.view_as(x), x_lv.view_as(x)                   # :100-101
        return self._finish(x, mu_idx, num_segs, (z1_mu, z1_lv, z1_s), (z2_mu, z2_lv, z2_s),
                            (x_mu, x_lv), mu2)


# --------------------------------------------------------------------------------------
# LSTM FHVAE restatement (O3) -- parity unpinned by the reference (fhvae.py is a stub).
# --------------------------------------------------------------------------------------
class _LSTMPre(nn.Module):
    def __init__(self, in_dim, hus):
        super().__init__()
        assert len(set(hus)) == 1, "all layers of one LSTM stack share a width"
        self.lstm = nn.LSTM(in_dim, hus[0], num_layers=len(hus), batch_first=True)


class FHVAEOracle(_FHVAEBase):
    """z2-enc: LSTM over x, concat of final h of all layers -> Gaussian.
    z1-enc: LSTM over cat[x_t, z2_sample] -> same.  decoder: LSTM fed cat[z1_s, z2_s] at every t,
    all T outputs of the last layer -> per-frame Gaussian heads.  PyTorch nn.LSTM conventions
    (gate order i,f,g,o; b_ih + b_hh; zero initial state)."""

    def __init__(self, input_size, z1_hus=(256, 256), z2_hus=(256, 256), z1_dim=32, z2_dim=32,
                 x_hus=(256, 256), *, seg_len=20, num_seqs=1000, detach_px=False, prior_grad=True,
                 ref_log_qy=False):
        super().__init__()
        self.model = "fhvae"
        self.pz1 = [0.0, np.float32(PZ1[1])]
        self.pmu2 = [0.0, np.float32(PMU2[1])]
        self.z1_hus, self.z2_hus, self.x_hus = list(z1_hus), list(z2_hus), list(x_hus)
        self.z1_dim, self.z2_dim = z1_dim, z2_dim
        self.detach_px, self.prior_grad, self.ref_log_qy = detach_px, prior_grad, ref_log_qy
        assert input_size % seg_len == 0
        self.seg_len, self.feat_dim = seg_len, input_size // seg_len
        F = self.feat_dim
        self.z1_pre_encoder = _LSTMPre(F + z2_dim, self.z1_hus)
        self.z2_pre_encoder = _LSTMPre(F, self.z2_hus)
        self.z1_gauss_layer = _Gauss(sum(self.z1_hus), z1_dim)
        self.z2_gauss_layer = _Gauss(sum(self.z2_hus), z2_dim)
        self.pre_decoder = _LSTMPre(z1_dim + z2_dim, self.x_hus)
        self.dec_gauss_layer = _Gauss(self.x_hus[-1], F)
        self.mu2_table = nn.Parameter(torch.randn(num_seqs, z2_dim))

    @staticmethod
    def _final_h(hn):                       # (L, B, H) -> (B, L*H), layer 0 first
        return hn.transpose(0, 1).reshape(hn.shape[1], -1)

    def forward(self, x, mu_idx, num_seqs, num_segs, eps=None):
        B, T, _ = x.shape
        mu2 = self.mu2_table[mu_idx]
        _, (hn, _) = self.z2_pre_encoder.lstm(x)
        z2_mu, z2_lv = self.z2_gauss_layer(self._final_h(hn))
        z2_s = reparam(z2_mu, z2_lv, _draw_eps(eps, "z2", z2_mu))
        _, (hn, _) = self.z1_pre_encoder.lstm(torch.cat([x, z2_s.unsqueeze(1).expand(B, T, -1)], -1))
        z1_mu, z1_lv = self.z1_gauss_layer(self._final_h(hn))
        z1_s = reparam(z1_mu, z1_lv, _draw_eps(eps, "z1", z1_mu))
        dec_in = torch.cat([z1_s, z2_s], -1).unsqueeze(1).expand(B, T, -1)
        out, _ = self.pre_decoder.lstm(dec_in)
        x_mu, x_lv = self.dec_gauss_layer(out)
        return self._finish(x, mu_idx, num_segs, (z1_mu, z1_lv, z1_s), (z2_mu, z2_lv, z2_s),
                            (x_mu, x_lv), mu2)


# --------------------------------------------------------------------------------------
# Second, independent pin of O3: a hand-written fp64 numpy LSTM cell + closed-form loss (SURVEY.md
# Appendix C), sharing NO code with torch.nn.LSTM / autograd.  tests/test_oracle.py checks FHVAEOracle
# against it, so a change of torch's LSTM conventions (gate order i,f,g,o; b_ih + b_hh; zero initial
# state; layer stacking) cannot silently move the oracle.
# --------------------------------------------------------------------------------------
def _sigm(a):
    return 1.0 / (1.0 + np.exp(-a))


def lstm_stack_fp64(x, weights):
    """x (B,T,In) float64; weights = [(W_ih (4H,In_l), W_hh (4H,H), b_ih, b_hh)] per layer.
    Returns (outputs of the last layer (B,T,H), final h of every layer [(B,H)], cache for the backward)."""
    B, T, _ = x.shape
    inp, finals, cache = x, [], []
    for (W_ih, W_hh, b_ih, b_hh) in weights:
        H = W_hh.shape[1]
        h, c = np.zeros((B, H)), np.zeros((B, H))
        outs, steps = np.zeros((B, T, H)), []
        for t in range(T):
            g = inp[:, t] @ W_ih.T + b_ih + h @ W_hh.T + b_hh
            i, f, gg, o = _sigm(g[:, :H]), _sigm(g[:, H:2 * H]), np.tanh(g[:, 2 * H:3 * H]), _sigm(g[:, 3 * H:])
            c_new = f * c + i * gg
            h_new = o * np.tanh(c_new)
            steps.append((inp[:, t], h, c, i, f, gg, o, c_new))
            h, c = h_new, c_new
            outs[:, t] = h
        finals.append(h)
        cache.append(steps)
        inp = outs
    return inp, finals, cache


def lstm_stack_bwd_fp64(d_out, d_finals, weights, cache):
    """Closed-form BPTT (Appendix C).  d_out (B,T,H) gradient of the last layer's outputs, d_finals[l] (B,H)
    gradient of layer l's final h.  Returns (dx (B,T,In), [(dW_ih, dW_hh, db_ih, db_hh)] per layer)."""
    grads = [None] * len(weights)
    d_seq = d_out
    for l in reversed(range(len(weights))):
        W_ih, W_hh, _, _ = weights[l]
        steps = cache[l]
        T, H = len(steps), W_hh.shape[1]
        B = steps[0][1].shape[0]
        dW_ih, dW_hh, db = np.zeros_like(W_ih), np.zeros_like(W_hh), np.zeros(4 * H)
        dx = np.zeros((B, T, W_ih.shape[1]))
        dh_next, dc_next = np.zeros((B, H)), np.zeros((B, H))
        for t in reversed(range(T)):
            x_t, h_prev, c_prev, i, f, gg, o, c_new = steps[t]
            dh = d_seq[:, t] + dh_next + (d_finals[l] if t == T - 1 else 0.0)
            tc = np.tanh(c_new)
            dc = dc_next + dh * o * (1 - tc * tc)
            dg = np.concatenate([dc * gg * i * (1 - i), dc * c_prev * f * (1 - f), dc * i * (1 - gg * gg),
                                 dh * tc * o * (1 - o)], axis=1)
            dW_ih += dg.T @ x_t
            dW_hh += dg.T @ h_prev
            db += dg.sum(0)
            dx[:, t] = dg @ W_ih
            dh_next = dg @ W_hh
            dc_next = dc * f
        grads[l] = (dW_ih, dW_hh, db, db.copy())
        d_seq = dx
    return d_seq, grads


def fhvae_forward_fp64(model: "FHVAEOracle", x, mu_idx, num_segs, eps, alpha=10.0):
    """The whole FHVAE forward (Appendix B architecture, Appendix C maths) in numpy float64 from the weights of
    an FHVAEOracle: returns dict of the six outputs + loss.  Independent of nn.LSTM and of the torch loss code."""
    sd = {k: v.detach().double().numpy() for k, v in model.state_dict().items()}
    x = x.double().numpy()
    idx = mu_idx.numpy()
    nseg = num_segs.double().numpy() if torch.is_tensor(num_segs) else np.full(x.shape[0], float(num_segs))
    B, T, F = x.shape

    def stack(prefix, L):
        return [(sd[f"{prefix}.lstm.weight_ih_l{l}"], sd[f"{prefix}.lstm.weight_hh_l{l}"],
                 sd[f"{prefix}.lstm.bias_ih_l{l}"], sd[f"{prefix}.lstm.bias_hh_l{l}"]) for l in range(L)]

    def head(prefix, h):
        return (h @ sd[prefix + ".mulayer.weight"].T + sd[prefix + ".mulayer.bias"],
                h @ sd[prefix + ".logvar_layer.weight"].T + sd[prefix + ".logvar_layer.bias"])

    table = sd["mu2_table"]
    mu2 = table[idx]
    _, fin, _ = lstm_stack_fp64(x, stack("z2_pre_encoder", len(model.z2_hus)))
    z2_mu, z2_lv = head("z2_gauss_layer", np.concatenate(fin, axis=1))
    z2_s = z2_mu + eps["z2"].double().numpy() * np.exp(0.5 * z2_lv)
    x1 = np.concatenate([x, np.repeat(z2_s[:, None, :], T, axis=1)], axis=2)
    _, fin, _ = lstm_stack_fp64(x1, stack("z1_pre_encoder", len(model.z1_hus)))
    z1_mu, z1_lv = head("z1_gauss_layer", np.concatenate(fin, axis=1))
    z1_s = z1_mu + eps["z1"].double().numpy() * np.exp(0.5 * z1_lv)
    din = np.repeat(np.concatenate([z1_s, z2_s], axis=1)[:, None, :], T, axis=1)
    out, _, _ = lstm_stack_fp64(din, stack("pre_decoder", len(model.x_hus)))
    x_mu, x_lv = head("dec_gauss_layer", out)
    lg = lambda v, mu, lv: -0.5 * (LOG_2PI + lv + (v - mu) ** 2 * np.exp(-lv))
    kl = lambda pm, pl, qm, c: -0.5 * (1 + pl - c - ((pm - qm) ** 2 + np.exp(pl)) * np.exp(-c))
    log_px = lg(x, x_mu, x_lv).sum(axis=(1, 2))
    nk1 = -kl(z1_mu, z1_lv, 0.0, PZ1[1]).sum(1)
    nk2 = -kl(z2_mu, z2_lv, mu2, PZ2_LOGVAR).sum(1)
    log_pmu2 = lg(mu2, PMU2[0], PMU2[1]).sum(1)
    lb = log_px + nk1 + nk2 + log_pmu2 / nseg
    s = -((z2_mu[:, None, :] - table[None]) ** 2).sum(-1) / (2 * math.exp(PZ2_LOGVAR))
    mx = s.max(1, keepdims=True)
    lse = mx[:, 0] + np.log(np.exp(s - mx).sum(1))
    log_qy = s[np.arange(B), idx] - lse
    return {"lower_bound": lb, "log_qy": log_qy, "log_px_z": log_px, "neg_kld_z1": nk1, "neg_kld_z2": nk2,
            "log_pmu2": log_pmu2, "loss": -np.mean(lb + alpha * log_qy)}


# --------------------------------------------------------------------------------------
# mu2 estimation (hierarchical-sampling cache refresh / inference), utils.py:45-60
# --------------------------------------------------------------------------------------
def estimate_mu2_dict(z2_mu_batches: Sequence[torch.Tensor], idx_batches: Sequence[torch.Tensor]):
    """The reference's python-dict loop (utils.py:49-60) over precomputed posterior means."""
    nseg_table: Dict[int, float] = defaultdict(float)
    z2_sum_table: Dict[int, torch.Tensor] = {}
    for z2, idxs in zip(z2_mu_batches, idx_batches):
        for _y, _z2 in zip(idxs.tolist(), z2):
            z2_sum_table[_y] = z2_sum_table.get(_y, 0.0) + _z2
            nseg_table[_y] += 1
    r = math.exp(PZ2_LOGVAR) / math.exp(PMU2[1])            # utils.py:58 -> 0.25
    return {y: z2_sum_table[y] / (nseg_table[y] + r) for y in nseg_table}


def estimate_mu2_table(z2_mu: torch.Tensor, idx: torch.Tensor, num_seqs: int):
    """Batched form: rows never seen stay 0.  Returns (table (K,Z) float64-accumulated, counts (K,))."""
    z = torch.zeros(num_seqs, z2_mu.shape[1], dtype=torch.float64)
    n = torch.zeros(num_seqs, dtype=torch.float64)
    z.index_add_(0, idx, z2_mu.double())
    n.index_add_(0, idx, torch.ones_like(idx, dtype=torch.float64))
    r = math.exp(PZ2_LOGVAR) / math.exp(PMU2[1])
    out = torch.where(n[:, None] > 0, z / (n[:, None] + r), torch.zeros_like(z))
    return out, n.long()


def hierarchical_sample(seqlist: Sequence, k: int, seed: int) -> np.ndarray:
    """train_model.py:426-428 on the legacy global numpy RNG, seeded: np.random.seed(seed);
    np.random.choice(seqlist, k, replace=False).  Local label of an utterance = its position."""
    st = np.random.RandomState(seed)
    return st.choice(np.asarray(seqlist), k, replace=False)


def segment_starts(length: int, seg_len: int = 20, seg_shift: int = 8) -> np.ndarray:
    """datasets.py:176-181 (fixed shift)."""
    nseg = (length - seg_len) // seg_shift + 1
    return np.arange(max(nseg, 0)) * seg_shift


# --------------------------------------------------------------------------------------
# one training step, the loop body of train_model.py:446-454
# --------------------------------------------------------------------------------------
def make_adam(params, lr=1e-3, betas=(0.95, 0.999)):
    """train_model.py:409-411."""
    return torch.optim.Adam(params, lr=lr, betas=betas)


def train_step(model, optimizer, x, mu_idx, num_seqs, num_segs, alpha=10.0, eps=None):
    optimizer.zero_grad()
    out = model(x, mu_idx, num_seqs, num_segs, eps=eps)
    loss = loss_function(out[0], out[1], alpha)
    loss.backward()
    optimizer.step()
    return loss.detach(), out
