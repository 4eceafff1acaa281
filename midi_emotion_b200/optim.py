"""Optimiser step of the reference's training loop (train.py:319-325) on the device in a handful of launches:

    scaler.unscale_(optimizer); clip_grad_norm_(model.parameters(), clip); scaler.step(optimizer)       # reference
    total_norm = optimizer.step(grad_scale=s)            # ClipAdam(params, lr, max_grad_norm=clip)        # here

`ClipAdam` is a `torch.optim.Optimizer` with torch.optim.Adam's hyper-parameters, state names (`step`, `exp_avg`,
`exp_avg_sq`) and `state_dict()` layout, so `optimizer.pt` files (train.py:180-182,401) move both ways.  The
arithmetic runs in `me_grad_sqnorm_partials` / `me_adam_prepare` / `me_adam_update` (csrc/optimizer.cu); there
is no host synchronisation: the clip coefficient, the skip-on-non-finite decision that GradScaler.step makes
(train.py:323) and the step count live on the device.  CUDA only, fp32 parameters, no fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib


class ClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 max_grad_norm: Optional[float] = None):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("ClipAdam: invalid hyper-parameters")
        # the key set of torch.optim.Adam's param_groups, so that state dicts are interchangeable
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None,
                        decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self.max_grad_norm = max_grad_norm
        self._scratch: Optional[torch.Tensor] = None
        self._stats: List[torch.Tensor] = []
        self._table = None          # (pointer key, host-side me_adam_tensor array over all groups)

    # ------------------------------------------------------------------ state
    def _group_step(self, group, params, device) -> torch.Tensor:
        """One device-resident float32 step counter per parameter group, shared by the `step` entry of every
        parameter of the group (torch.optim.Adam keeps one such tensor per parameter; they all hold the same
        number).  Created from the largest `step` found in the state (0 for a fresh optimiser; the loaded value
        after load_state_dict -- a one-time host read)."""
        shared = group.get("_step_dev")
        if shared is None or shared.device != device:
            start = 0.0
            for p in params:
                st = self.state[p]
                if "step" in st:
                    start = max(start, float(st["step"]))
            shared = torch.full((1,), start, dtype=torch.float32, device=device)
            group["_step_dev"] = shared
            group["_step_view"] = shared.view(())
        view = group["_step_view"]
        for p in params:
            st = self.state[p]
            if st.get("step") is not view:
                st["step"] = view
        return shared

    def state_dict(self):
        sd = super().state_dict()
        groups = [{k: v for k, v in g.items() if k not in ("_step_dev", "_step_view")} for g in sd["param_groups"]]
        state = {}
        for k, st in sd["state"].items():
            st = dict(st)
            if "step" in st:                       # torch.optim.Adam's default layout: a 0-d float32 CPU tensor
                st["step"] = st["step"].detach().to("cpu", torch.float32).reshape(()).clone()
            state[k] = st
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for g in self.param_groups:                # counters are rebuilt from the loaded `step` entries
            g.pop("_step_dev", None)
            g.pop("_step_view", None)
        self._table = None

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        """Applies one update; returns the total gradient norm (0-d device tensor, the value clip_grad_norm_
        returns; zero when neither clipping nor a gradient scale asked for it; a view of `last_stats`, valid until
        the next step -- clone it to keep it).  `grad_scale`: the gradients are
        divided by it first (GradScaler.unscale_).  When a norm is taken, a non-finite norm skips the update and
        leaves the step count unchanged, as GradScaler.step does."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        groups = []
        for group in self.param_groups:
            if group.get("amsgrad") or group.get("maximize") or group.get("decoupled_weight_decay"):
                raise RuntimeError("ClipAdam: amsgrad / maximize / decoupled_weight_decay are not implemented")
            ps = [p for p in group["params"] if p.grad is not None]
            if ps:
                groups.append((group, ps))
        if not groups:
            return loss
        device = groups[0][1][0].device
        for _, ps in groups:
            for p in ps:
                if not p.is_cuda or p.device != device:
                    raise RuntimeError("midi_emotion_b200: ClipAdam needs CUDA parameters on one device (no CPU fallback)")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("midi_emotion_b200: ClipAdam needs dense float32 parameters and gradients")
                if not p.is_contiguous():
                    raise RuntimeError("midi_emotion_b200: ClipAdam needs contiguous parameters")
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                st = self.state[p]
                if "exp_avg" not in st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        stream = torch.cuda.current_stream(device).cuda_stream

        # one host-side descriptor array over all groups (pointers travel as kernel parameters): parameter and
        # state pointers are stable, gradient pointers move from step to step (zero_grad(set_to_none=True))
        key = tuple((id(g), tuple(p.data_ptr() for p in ps), tuple(self.state[p]["exp_avg"].data_ptr() for p in ps))
                    for g, ps in groups)
        if self._table is None or self._table[0] != key:
            flat = (_lib.AdamTensor * sum(len(ps) for _, ps in groups))()
            k = 0
            for _, ps in groups:
                for p in ps:
                    st = self.state[p]
                    flat[k].param, flat[k].numel = p.data_ptr(), p.numel()
                    flat[k].exp_avg, flat[k].exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                    k += 1
            self._table = (key, flat)
        flat = self._table[1]
        k = 0
        for _, ps in groups:
            for p in ps:
                flat[k].grad = p.grad.data_ptr()
                k += 1
        n_all = k

        want_norm = self.max_grad_norm is not None or grad_scale != 1.0
        n_partials, partials_ptr = 0, None
        if want_norm:                                  # one global norm over every group (train.py:321-322)
            n_partials = int(_lib.load().me_grad_sqnorm_chunks(flat, n_all))
            if n_partials < 0:
                raise RuntimeError("midi_emotion_b200: bad tensor list")
            if self._scratch is None or self._scratch.numel() < n_partials or self._scratch.device != device:
                self._scratch = torch.empty(max(n_partials, 1), dtype=torch.float32, device=device)
            partials_ptr = self._scratch.data_ptr()
            _lib.call("me_grad_sqnorm_partials", flat, n_all, float(grad_scale), partials_ptr, self._scratch.numel(),
                      stream)
        while len(self._stats) < len(groups):
            self._stats.append(torch.zeros(8, dtype=torch.float32, device=device))
        first = 0
        for gi, (group, ps) in enumerate(groups):
            if self._stats[gi].device != device:
                self._stats[gi] = torch.zeros(8, dtype=torch.float32, device=device)
            stats = self._stats[gi]
            step_dev = self._group_step(group, ps, device)
            b1, b2 = group["betas"]
            sub = C.cast(C.byref(flat, first * C.sizeof(_lib.AdamTensor)), C.POINTER(_lib.AdamTensor))
            _lib.call("me_adam_prepare", partials_ptr, n_partials, float(self.max_grad_norm or 0.0), float(group["lr"]),
                      float(b1), float(b2), step_dev.data_ptr(), stats.data_ptr(), stream)
            _lib.call("me_adam_update", sub, len(ps), float(b1), float(b2), float(group["eps"]),
                      float(group["weight_decay"]), float(grad_scale), stats.data_ptr(), stream)
            first += len(ps)
        return loss if closure is not None else self._stats[0][0]

    @property
    def last_stats(self) -> torch.Tensor:
        """Device f32[8] of the last step (first group): total_norm, clip coefficient, skipped flag, step count,
        lr / bias_correction1, sqrt(bias_correction2)."""
        return self._stats[0]
