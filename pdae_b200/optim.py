"""Fused multi-tensor Adam + EMA.

Replaces, for the trainable parameters of the hot path, the pair the reference runs every step:
  * `torch.optim.Adam([...5 groups...], lr, betas, eps, weight_decay)` and `scaler.step(optimizer)`
    (trainer/train_representation_learning.py:54-69, :118-119);
  * `accumulate(decay)` -- a python loop over every named parameter doing
    `ema.data.mul_(decay).add_(p.data, alpha=1-decay)` (trainer/train_representation_learning.py:192-212).
One `pdae_adam_ema_step` launch per parameter group updates p, exp_avg, exp_avg_sq and the EMA copy in a single pass
over HBM (reads p,g,m,v,ema; writes p,m,v,ema: 36 B/element, versus ~100 B/element for the unfused sequence).
State layout and `state_dict()` keys follow torch.optim.Adam (`step`, `exp_avg`, `exp_avg_sq`), so checkpoints written by
the reference trainer (`'optimizer'` entry) load unchanged.  No CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np
import torch

from . import _native

_CHUNK = 65536


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


class FusedAdamEMA(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 ema_decay: float = 0.9999, ema_every: int = 1):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FusedAdamEMA: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.ema_decay, self.ema_every = float(ema_decay), int(ema_every)
        self._ema: Dict[torch.Tensor, torch.Tensor] = {}
        self._steps = 0
        self._maps: Dict[tuple, torch.Tensor] = {}

    # -- EMA pairing, by name like the reference's accumulate() ---------------------------------------------------------
    def attach_ema(self, model: torch.nn.Module, ema_model: torch.nn.Module) -> None:
        """Pair every trainable parameter of `model` with the same-named parameter of `ema_model`
        (train_representation_learning.py:196-212: only `requires_grad` parameters are accumulated)."""
        ema_named = dict(ema_model.named_parameters())
        for k, p in model.named_parameters():
            if p.requires_grad:
                e = ema_named[k]
                if e.shape != p.shape or e.dtype != torch.float32 or not e.is_contiguous():
                    raise ValueError(f"attach_ema: {k}: EMA copy must be a contiguous fp32 tensor of the same shape")
                self._ema[p] = e

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        """One optimizer step (+ EMA when `steps % ema_every == 0`).  `grad_scale` multiplies every gradient first:
        1/world_size after a sum all-reduce, or 1/loss_scale in place of GradScaler.unscale_."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._steps += 1
        do_ema = self.ema_every > 0 and self._steps % self.ema_every == 0
        L = _native.lib()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            if dev.type != "cuda":
                raise _native.NativeError("FusedAdamEMA.step: CUDA parameters required (no CPU fallback)")
            rows = np.empty((len(ps), 6), dtype=np.int64)
            step_no = None
            for i, p in enumerate(ps):
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise ValueError("FusedAdamEMA: parameters must be contiguous fp32")
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = p.grad = g.float().contiguous()
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                s = int(st["step"].item())
                if step_no is None:
                    step_no = s
                elif s != step_no:
                    raise ValueError("FusedAdamEMA: parameters of one group must share a step count")
                e = self._ema.get(p) if do_ema else None
                rows[i] = (_ptr(p), _ptr(g), _ptr(st["exp_avg"]), _ptr(st["exp_avg_sq"]), _ptr(e), p.numel())
            key = (gi, tuple(int(r[5]) for r in rows))
            bmap = self._maps.get(key)
            if bmap is None:
                pairs = [(i, c) for i, r in enumerate(rows) for c in range((int(r[5]) + _CHUNK - 1) // _CHUNK)]
                bmap = self._maps[key] = torch.tensor(pairs, dtype=torch.int32, device=dev).contiguous()
            table = torch.from_numpy(rows).pin_memory().to(dev, non_blocking=True)
            b1, b2 = group["betas"]
            rc = L.pdae_adam_ema_step(table.data_ptr(), bmap.data_ptr(), bmap.shape[0], _CHUNK, float(group["lr"]), float(b1),
                                      float(b2), float(group["eps"]), float(group["weight_decay"]), step_no, float(grad_scale),
                                      self.ema_decay if do_ema else -1.0,
                                      ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            _native.check(rc, "pdae_adam_ema_step")
            # the kernel wrote p and the EMA copies through raw pointers: bump their autograd version counters so every
            # packed (re-laid-out / bf16) weight copy keyed on (storage, version) is re-derived on its next use
            touched = list(ps) + ([self._ema[p] for p in ps if p in self._ema] if do_ema else [])
            torch.autograd.graph.increment_version(touched)
        return loss
