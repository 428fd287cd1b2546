"""``FusedAdam``: drop-in for ``torch.optim.Adam(model.parameters(), lr, betas)`` as constructed at
train_model.py:409-411, running one flat multi-tensor kernel (``fhvae_adam_flat``) per model instead
of the reference's per-tensor foreach update (train_model.py:454)."""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib
from .plan import current_stream_ptr, ptr


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam runs ONE flat kernel per model: a single param group (model.parameters(), "
                             "train_model.py:409-411) is required")
        self.grad_scale = float(grad_scale)
        self._flat_state: Dict[int, dict] = {}
        self._modules = []
        seen = set()
        for group in self.param_groups:
            for p in group["params"]:
                owner = getattr(p, "_fhvae", None)
                if owner is None:
                    raise ValueError("FusedAdam only drives parameters of pytorch_scalablefhvae_b200 modules")
                mod = owner[0]()
                if id(mod) not in seen:
                    seen.add(id(mod))
                    self._modules.append(mod)
        for mod in self._modules:
            if len([p for g in self.param_groups for p in g["params"] if p._fhvae[0]() is mod]) != len(mod._plist):
                raise ValueError("FusedAdam needs all parameters of a model (model.parameters())")

    def _hyper(self):
        g = self.param_groups[0]
        if g.get("weight_decay", 0) or g.get("amsgrad", False) or g.get("maximize", False):
            raise ValueError("FusedAdam implements plain Adam (train_model.py:409-411): weight_decay / amsgrad / "
                             "maximize are not supported")
        return float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"])

    def hyper_key(self):
        """What a captured step bakes in (model.train_step re-captures its graph when this changes)."""
        return self._hyper() + (float(self.grad_scale),)

    def _state_for(self, mod):
        flat = mod._ensure_flat()
        st = self._flat_state.get(id(mod))
        if st is None or st["m"].device != flat.device:
            old = st
            st = {"m": torch.zeros_like(flat), "v": torch.zeros_like(flat),
                  "step": torch.zeros(1, dtype=torch.int32, device=flat.device),
                  "done": torch.zeros(1, dtype=torch.int32, device=flat.device)}
            if old is not None:
                st["m"].copy_(old["m"]); st["v"].copy_(old["v"]); st["step"].copy_(old["step"])
            self._flat_state[id(mod)] = st
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lr, b1, b2, eps = self._hyper()
        for mod in self._modules:
            self.step_flat(mod, mod.packed_grads(), lr, b1, b2, eps)
        return loss

    def step_flat(self, mod, gflat, lr=None, b1=None, b2=None, eps=None):
        """Adam on the model's flat parameter buffer given a flat gradient buffer (graph-capturable)."""
        h = self._hyper()
        lr = h[0] if lr is None else lr
        b1 = h[1] if b1 is None else b1
        b2 = h[2] if b2 is None else b2
        eps = h[3] if eps is None else eps
        st = self._state_for(mod)
        flat = mod._flat
        rc = _lib.fn("fhvae_adam_flat")(ptr(flat), ptr(gflat), ptr(st["m"]), ptr(st["v"]), flat.numel(),
                                        lr, b1, b2, eps, self.grad_scale, ptr(st["step"]), ptr(st["done"]),
                                        current_stream_ptr())
        _lib.check(rc, "fhvae_adam_flat")

    @torch.no_grad()
    def reset_state(self, mod, name: str):
        """Zero the Adam moments of one parameter (hierarchical sampling: the rows of the active mu2 table stand
        for NEW utterances after every re-sample, train_model.py:424-436, so their moments must not carry over)."""
        st = self._state_for(mod)
        o, n = mod._off[name], mod._shape[name]
        cnt = 1
        for d in n:
            cnt *= int(d)
        st["m"][o:o + cnt].zero_()
        st["v"][o:o + cnt].zero_()

    # ---- torch.optim.Adam-compatible state (utils.py:142 saves optimizer.state_dict())
    def state_dict(self):
        state, idx = {}, 0
        groups = []
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                mod, name = p._fhvae[0](), p._fhvae[1]
                st = self._flat_state.get(id(mod))
                if st is not None:
                    o, n = mod._off[name], p.numel()
                    state[idx] = {"step": st["step"].float().cpu().reshape(()),
                                  "exp_avg": st["m"][o:o + n].view_as(p).clone(),
                                  "exp_avg_sq": st["v"][o:o + n].view_as(p).clone()}
                ids.append(idx)
                idx += 1
            groups.append({**{k: v for k, v in g.items() if k != "params"}, "params": ids})
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        idx = 0
        for g, sg in zip(self.param_groups, sd["param_groups"]):
            for k, v in sg.items():
                if k != "params":
                    g[k] = v
            for p in g["params"]:
                mod, name = p._fhvae[0](), p._fhvae[1]
                st = self._state_for(mod)
                ps = sd["state"].get(idx)
                if ps is not None:
                    o, n = mod._off[name], p.numel()
                    st["m"][o:o + n].view_as(p).copy_(ps["exp_avg"])
                    st["v"][o:o + n].view_as(p).copy_(ps["exp_avg_sq"])
                    st["step"].fill_(int(ps["step"]))
                idx += 1

    def steps_taken(self, mod=None) -> int:
        mod = mod or self._modules[0]
        return int(self._state_for(mod)["step"].item())
