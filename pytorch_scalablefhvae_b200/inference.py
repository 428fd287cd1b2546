"""Posterior extraction over whole utterances (BASELINE config 4) -- what the reference's eval_model.py:55-59
leaves as three TODOs.  Utterances stay packed in HBM; 20-frame segments at `seg_shift` stride
(datasets.py:155-185) are cut by a device-side gather kernel, pushed through the encoders only, and the
per-utterance mu2 follows utils.estimate_mu2_dict (utils.py:45-60) as a deterministic segmented sum.
Utterances are independent: for N GPUs shard the utterance list by rank -- no collective is needed.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .plan import current_stream_ptr, ptr

R_MU2 = 0.25 / 1.0     # exp(pz2_logvar) / exp(pmu2_logvar), utils.py:58


def segment_table(lengths: Sequence[int], seg_len: int = 20, seg_shift: int = 8):
    """(start_row (S,), utt_id (S,), nsegs (U,)) for utterances packed back to back; datasets.py:176-181."""
    starts, utts, nsegs = [], [], []
    off = 0
    for u, l in enumerate(lengths):
        n = max((int(l) - seg_len) // seg_shift + 1, 0)
        nsegs.append(n)
        if n:
            starts.append(off + np.arange(n, dtype=np.int64) * seg_shift)
            utts.append(np.full(n, u, dtype=np.int64))
        off += int(l)
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.int64)
    return cat(starts), cat(utts), np.asarray(nsegs, dtype=np.int64)


def default_batch_size(model, target: int = 2048) -> int:
    """Segments per forward-only batch: the tensor-core recurrence serves ``fhvae_lstm_wave_rows_per_launch`` rows per
    launch (288 for 2x256 on 148 SMs) and runs a larger batch as consecutive launches, so a batch of 2048 = 7 full
    launches + one with a single 32-row group; 2016 = 7 x 288 does the same work in 7."""
    hus = [getattr(model, a, None) for a in ("z2_hus", "z1_hus")]
    if (getattr(model, "model", "") == "fhvae" and all(h is not None and len(h) == 2 and len(set(h)) == 1 for h in hus)
            and hus[0][0] == hus[1][0]):
        rows = _lib.fn("fhvae_lstm_wave_rows_per_launch")(int(hus[0][0]), 2)
        if rows > 0:
            return max(rows, target // rows * rows)
    return target


@torch.no_grad()
def extract_posteriors(model, feats: torch.Tensor, lengths: Sequence[int], seg_shift: int = 8,
                       batch_size: Optional[int] = None, mean: Optional[torch.Tensor] = None,
                       std: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """feats: (sum(lengths), F) fp32 on the model's device (all utterances packed).  Returns z1_mu (S,Z1),
    z2_mu (S,Z2) per segment, mu2 (U,Z2) per utterance, plus seg_utt (S,) and nsegs (U,).  ``batch_size=None``:
    ~2048 segments, rounded to a whole number of recurrence launches (``default_batch_size``)."""
    if batch_size is None:
        batch_size = default_batch_size(model)
    dev = feats.device
    if not feats.is_cuda:
        raise RuntimeError("extract_posteriors needs the packed features on the GPU (no CPU path)")
    T = model.seg_len if hasattr(model, "seg_len") else model.input_size // feats.shape[1]
    F = feats.shape[1]
    starts_np, utt_np, nsegs_np = segment_table(lengths, T, seg_shift)
    S, U = len(starts_np), len(lengths)
    starts, seg_utt = torch.from_numpy(starts_np).to(dev), torch.from_numpy(utt_np).to(dev)
    Z1, Z2 = model.z1_dim, model.z2_dim
    z1_mu, z2_mu = torch.empty(S, Z1, device=dev), torch.empty(S, Z2, device=dev)
    zsum, cnt = torch.zeros(U, Z2, device=dev), torch.zeros(U, device=dev)
    inv_std = (1.0 / std).contiguous() if std is not None else None
    gather = _lib.fn("fhvae_gather_segments")
    accumulate = _lib.fn("fhvae_mu2_accumulate")
    xb = torch.empty(batch_size, T, F, device=dev)
    for s0 in range(0, S, batch_size):
        nb = min(batch_size, S - s0)
        x = xb if nb == batch_size else torch.empty(nb, T, F, device=dev)
        _lib.check(gather(ptr(feats), ptr(starts, s0), ptr(mean) if mean is not None else None,
                          ptr(inv_std) if inv_std is not None else None, ptr(x), nb, T, F, feats.shape[0],
                          current_stream_ptr()), "fhvae_gather_segments")
        enc = model.encode(x)
        z1_mu[s0:s0 + nb].copy_(enc["z1_mu"])
        z2_mu[s0:s0 + nb].copy_(enc["z2_mu"])
        z2h = enc["z2_mu"]                                  # view of the (nb, 2*Z2) head: leading dim 2*Z2
        _lib.check(accumulate(ptr(z2h), 2 * Z2, ptr(seg_utt, s0), ptr(zsum), ptr(cnt), nb, Z2, U, None,
                              current_stream_ptr()), "fhvae_mu2_accumulate")
    mu2 = torch.zeros(U, Z2, device=dev)
    _lib.check(_lib.fn("fhvae_mu2_estimate_finish")(ptr(zsum), ptr(cnt), ptr(mu2), R_MU2, U, Z2,
                                                    current_stream_ptr()), "fhvae_mu2_estimate_finish")
    return {"z1_mu": z1_mu, "z2_mu": z2_mu, "mu2": mu2, "seg_utt": seg_utt,
            "nsegs": torch.from_numpy(nsegs_np).to(dev)}



def shard_utterances(num_utts: int, rank: int, world: int) -> np.ndarray:
    """Utterances of `rank`: a contiguous block (utterances are independent: no collective on the data path)."""
    per = (num_utts + world - 1) // world
    return np.arange(rank * per, min(num_utts, (rank + 1) * per))


@torch.no_grad()
def extract_posteriors_sharded(model, feats_of, lengths: Sequence[int], rank: int, world: int, **kw):
    """BASELINE config 4 at N GPUs: every rank extracts the posteriors of ITS block of utterances
    (shard_utterances); ``feats_of(utt_ids) -> (sum(len), F)`` CUDA features of those utterances, packed.
    Returns extract_posteriors' dict for the local block plus ``utts`` (their global ids).  Gathering the per-rank
    results (if wanted at all -- eval_model.py writes per-utterance files) is left to the caller."""
    mine = shard_utterances(len(lengths), rank, world)
    lens = np.asarray(lengths)[mine]
    out = extract_posteriors(model, feats_of(mine), lens, **kw)
    out["utts"] = torch.from_numpy(mine)
    return out
