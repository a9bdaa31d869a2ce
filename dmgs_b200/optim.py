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
            for p in g["params"]:
                grad = grads.get(g.get("name")) if grads is not None else p.grad
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

    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is not None:
                    if set_to_none:
                        p.grad = None
                    else:
                        p.grad.zero_()
