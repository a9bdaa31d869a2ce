"""Fused multi-tensor Adam on the device (libdmgs_raster.so: dmgs_adam_step).

Host-side mirror of how the reference drives ``torch.optim.Adam(l, lr=0.0, eps=1e-15)`` with one
parameter group per tensor (scene/gaussian_geo_model_finetune.py:526-537; step + zero_grad at
train_geo_stage3.py:164-166): ``FusedAdam(param_groups, lr=0.0, eps=1e-15)`` takes the same list of
``{'params': [tensor], 'lr': ..., 'name': ...}`` dicts, exposes ``param_groups`` (so the reference's
``update_learning_rate`` loop that assigns ``param_group['lr']`` works unchanged) and ``state``.

Differences that make it B200-native: ``step()`` is ONE kernel launch over every group (up to 8 per
launch), gradients may live in a view-batched ``FlatGradBuffer`` (``step(grads=buf.views)``), the view
averaging (``grad_scale``) and ``zero_grad`` are folded into the same pass.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional

import torch

from . import _lib as L


class FusedAdam:
    def __init__(self, params: Iterable[dict], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps)
        self.param_groups = []
        for g in params:
            g = dict(g)
            ps = g["params"]
            g["params"] = [ps] if isinstance(ps, torch.Tensor) else list(ps)
            for k, v in self.defaults.items():
                g.setdefault(k, v)
            for p in g["params"]:
                if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters; there is no CPU path")
            self.param_groups.append(g)
        self.state: Dict[torch.Tensor, dict] = {}

    def _state(self, p):
        st = self.state.get(p)
        if st is None:
            st = {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
            self.state[p] = st
        return st

    @torch.no_grad()
    def step(self, grads: Optional[Dict[str, torch.Tensor]] = None, grad_scale: float = 1.0, zero_grad: bool = False):
        """grads: optional {group name: gradient tensor} (e.g. FlatGradBuffer.views); default = p.grad.
        Groups may carry 'lr_hi', 'period', 'split': elements with (i % period) >= split use lr_hi."""
        lib = L.lib()
        segs, keep = [], []
        betas, eps = None, None
        for g in self.param_groups:
            if grads is not None:
                # gradients are looked up by GROUP name: that is only unambiguous for one tensor per group (the
                # reference's layout, finetune.py:526-537), and a name missing from `grads` is an error, not a skip
                if len(g["params"]) != 1:
                    raise RuntimeError(f"FusedAdam.step(grads=...): group {g.get('name')!r} holds {len(g['params'])} tensors; "
                                       "name-keyed gradients need one tensor per group")
                if g.get("name") not in grads:
                    raise KeyError(f"FusedAdam.step(grads=...): no gradient for group {g.get('name')!r} "
                                   f"(have {sorted(grads)})")
            for p in g["params"]:
                grad = grads[g.get("name")] if grads is not None else p.grad
                if grad is None:
                    continue
                if grad.dtype != torch.float32 or not grad.is_contiguous() or grad.numel() != p.numel():
                    raise RuntimeError(f"gradient of group {g.get('name')!r} must be contiguous fp32 of the parameter's size")
                if betas is None:
                    betas, eps = tuple(g["betas"]), float(g["eps"])
                elif betas != tuple(g["betas"]) or eps != float(g["eps"]):
                    raise RuntimeError("FusedAdam: all groups of one step share betas and eps")
                st = self._state(p)
                st["step"] += 1
                segs.append((st["step"], L.AdamSegment(p.data_ptr(), grad.data_ptr(), st["exp_avg"].data_ptr(),
                                                       st["exp_avg_sq"].data_ptr(), p.numel(), float(g["lr"]),
                                                       float(g.get("lr_hi", g["lr"])), int(g.get("period", 0)),
                                                       int(g.get("split", 0)))))
                keep.append(grad)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        # one launch per distinct step count (normally one), at most DMGS_ADAM_MAX_SEGMENTS tensors each
        for t in sorted({s for s, _ in segs}):
            group = [seg for s, seg in segs if s == t]
            for o in range(0, len(group), 8):
                chunk = group[o:o + 8]
                arr = (L.AdamSegment * len(chunk))(*chunk)
                L.check(lib.dmgs_adam_step(len(chunk), arr, betas[0], betas[1], eps, t, float(grad_scale),
                                           int(zero_grad), stream), "dmgs_adam_step")

    # ---- torch.optim.Optimizer's checkpoint surface (the reference's capture() calls optimizer.state_dict(),
    # scene/gaussian_geo_model_mlp_flex.py:102, finetune.py:92-107; restore loads it back)
    def state_dict(self) -> dict:
        """Same layout as torch.optim.Optimizer.state_dict(): {'state': {index: {...}}, 'param_groups': [...]}
        with parameters replaced by their running index."""
        index, packed_groups = {}, []
        for g in self.param_groups:
            pg = {k: v for k, v in g.items() if k != "params"}
            pg["params"] = []
            for p in g["params"]:
                index.setdefault(id(p), len(index))
                pg["params"].append(index[id(p)])
            packed_groups.append(pg)
        state = {}
        for p, st in self.state.items():
            if id(p) in index:
                state[index[id(p)]] = {"step": torch.tensor(float(st["step"])), "exp_avg": st["exp_avg"],
                                       "exp_avg_sq": st["exp_avg_sq"]}
        return {"state": state, "param_groups": packed_groups}

    def load_state_dict(self, sd: dict):
        groups = sd["param_groups"]
        if len(groups) != len(self.param_groups) or any(len(a["params"]) != len(b["params"])
                                                       for a, b in zip(groups, self.param_groups)):
            raise ValueError("loaded state dict has different parameter groups")
        by_index = {}
        for saved, g in zip(groups, self.param_groups):
            for i, p in zip(saved["params"], g["params"]):
                by_index[i] = p
            for k, v in saved.items():
                if k != "params":
                    g[k] = v
        self.state = {}
        for i, st in sd["state"].items():
            p = by_index[int(i)]
            step = st["step"]
            self.state[p] = {"step": int(step.item() if isinstance(step, torch.Tensor) else step),
                             "exp_avg": st["exp_avg"].to(p.device, torch.float32).contiguous().clone(),
                             "exp_avg_sq": st["exp_avg_sq"].to(p.device, torch.float32).contiguous().clone()}

    def add_param_group(self, group: dict):
        g = dict(group)
        ps = g["params"]
        g["params"] = [ps] if isinstance(ps, torch.Tensor) else list(ps)
        for k, v in self.defaults.items():
            g.setdefault(k, v)
        for p in g["params"]:
            if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters; there is no CPU path")
            if any(p is q for og in self.param_groups for q in og["params"]):
                raise ValueError("some parameters appear in more than one parameter group")
        self.param_groups.append(g)

    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is not None:
                    if set_to_none:
                        p.grad = None
                    else:
                        p.grad.zero_()
