"""Checkpoint interop with the reference (utils.py:63-102 load, :116-152 save): same dict keys and
file names, so either side can read the other's files.  Fixes Appendix A13: `model_params` also carries
`input_size` (the reference's 5-tuple cannot re-construct the model) under the extra key
`model_params_full`; the reference's own 5-tuple is kept as is."""
from __future__ import annotations

import shutil
from pathlib import Path

import torch

from .model import FHVAE, SimpleFHVAE


def save_checkpoint(model, optimizer, summary_list, values_dict, run_info: str, epoch: int, best_epoch: int,
                    val_lower_bound: float, best_val_lb: float, checkpoint_dir: str) -> Path:
    checkpoint = {
        "best_val_lb": best_val_lb, "best_epoch": best_epoch, "epoch": epoch, "model_type": model.model,
        "model_params": (model.z1_hus, model.z2_hus, model.z1_dim, model.z2_dim, model.x_hus),     # utils.py:135-141
        "model_params_full": {"input_size": model.input_size, "num_seqs": model.mu2_table.shape[0],
                              "seg_len": getattr(model, "seg_len", None)},
        "optimizer": optimizer.state_dict(), "state_dict": model.state_dict(),
        "summary_vals": summary_list, "values": values_dict,
    }
    f_str = f"{model.model}_{run_info}_e{epoch}"
    f_path = Path(checkpoint_dir) / f"{f_str}.tar"
    torch.save(checkpoint, f_path)
    if best_epoch == epoch:
        shutil.copyfile(f_path, Path(checkpoint_dir) / f"best_model_{f_str}.tar")
    return f_path


def load_checkpoint_file(checkpoint_file, finetune: bool, input_size=None, **model_kw):
    """Same 6-tuple as utils.load_checkpoint_file (including its start_epoch = epoch + 2)."""
    ck = torch.load(checkpoint_file, map_location="cpu", weights_only=False)
    full = ck.get("model_params_full", {})
    input_size = input_size if input_size is not None else full.get("input_size")
    if input_size is None:
        raise ValueError("reference checkpoints do not store input_size (utils.py:135-141): pass input_size=")
    sd = ck["state_dict"]
    if "mu2_table" in sd:
        model_kw.setdefault("num_seqs", sd["mu2_table"].shape[0])
    cls = {"fhvae": FHVAE, "simple_fhvae": SimpleFHVAE}[ck["model_type"]]
    if cls is FHVAE and full.get("seg_len"):
        model_kw.setdefault("seg_len", full["seg_len"])
    model = cls(input_size, *ck["model_params"], **model_kw)
    # a reference checkpoint has no persistent table (Appendix A1): keep the fresh one in that case
    missing = model.load_state_dict(sd, strict="mu2_table" in sd)
    if not finetune:
        return model, ck["values"], ck["optimizer"], ck["epoch"] + 2, ck["best_val_lb"], ck["summary_vals"]
    return model, None, None, None, None, None
