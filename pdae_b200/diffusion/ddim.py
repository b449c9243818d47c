"""Respaced deterministic DDIM (reference surface: diffusion/ddim.py:7-207) on native kernels.

Same constructor (``DDIM(betas, timestep_map, device)``) and method names.  The per-step arithmetic
(23 ATen launches in the reference) is ONE fused kernel, ``pdae_ddim_step``; when the network is a
pdae_b200 ShiftUNet/UNet the loop drives the network's static plan buffers directly (no per-step
allocation, no host sync, no tqdm).
"""
from __future__ import annotations

import ctypes
import os
from functools import partial

import numpy as np
import torch

from .. import _native


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _f32(t):
    """The elementwise kernels read raw fp32 pointers: promote anything else (half tensors under AMP, double images from a
    dataset) the way the reference's type-promoting torch arithmetic would, instead of reinterpreting the bytes."""
    return None if t is None else t.to(torch.float32).contiguous()


def _t64(t, B, what):
    t = t.to(torch.int64).contiguous()
    if t.numel() != B:
        raise ValueError(f"{what}: t must hold one timestep per sample ({B}), got {tuple(t.shape)}")
    return t


class _StepRunner:
    """One network plan + the loop bookkeeping + the fused update of a DDIM loop direction, replayed as ONE CUDA graph per
    step.  The graph is cached on the plan, keyed by (DDIM object, direction, shift) -- those fix the table pointers the
    update kernel reads.  `seek(i)` sets the device-side step counter; `step()` is a single graph launch (or, with
    PDAE_NO_GRAPH=1, the same launch sequence issued eagerly)."""

    def __init__(self, ddim: "DDIM", plan, x_in, t_in, eps, grad, direction: str, C: int):
        self.d, self.plan, self.x_in, self.t_in, self.eps, self.grad, self.direction = ddim, plan, x_in, t_in, eps, grad, direction
        dev = ddim.device
        B = x_in.tensor.shape[0]
        self.B = B
        self.delta = -1 if direction == "sample" else 1
        # the update runs inside the step graph (separate kernel, or fused into the last head conv's epilogue); only a learn_sigma
        # head on the CUDA-core path (2C output channels, no fusable epilogue) needs it after the replay
        self.in_graph_update = eps.tensor.shape[1] == C or bool(getattr(plan, "head_fuse", None))
        self.fused = False
        self.C = C
        cache = plan.__dict__.setdefault("_step_cache", {})
        key = (id(ddim), direction, grad is not None)
        ent = cache.get(key)
        if ent is None:
            with torch.inference_mode(False):   # never inference tensors: the cache outlives the caller's autograd mode
                ent = {"counter": torch.zeros(1, dtype=torch.int64, device=dev),
                       "t_loc": torch.zeros(B, dtype=torch.int64, device=dev), "graph": None, "ddim": ddim}
            cache[key] = ent
        self.ent = ent

    def _fuse_target(self):
        """The image head that ends the step plan and can run the update in its epilogue (tensor-core heads only)."""
        hf = getattr(self.plan, "head_fuse", None) or {}
        if "grad" in hf:
            return hf["grad"], True
        if "eps" in hf:
            return hf["eps"], False
        return None, False

    def _launch_step(self):
        d, L = self.d, _native.lib()
        st = _stream(d.device)
        rc = L.pdae_ddim_select_t(_ptr(self.ent["counter"]), self.delta, _ptr(d.timestep_map), int(d.timestep_map.shape[0]),
                                  _ptr(self.ent["t_loc"]), _ptr(self.t_in.tensor), self.B, st)
        _native.check(rc, "pdae_ddim_select_t")
        self.plan._launch_all()
        if self.in_graph_update and not self.fused:
            x = self.x_in.tensor
            d._update(x, self.ent["t_loc"], self.eps.tensor, self.grad.tensor if self.grad is not None else None,
                      self.direction, out=x)

    def begin(self, use_graph: bool):
        self.plan.run_prologue()          # forced weight re-pack + step-invariant ops (label_emb(z), emb_z_layers)
        # DDIM update fused into the last head conv's epilogue: point its device-side descriptor at this loop's tables
        fuse, is_grad_head = self._fuse_target()
        self.fused = fuse is not None and self.in_graph_update
        self.fuse_buf = fuse if self.fused else None
        if self.fused:
            d = self.d
            tab = d.alphas_cumprod_prev if self.direction == "sample" else d.alphas_cumprod_next
            Ce = int(self.eps.tensor.shape[1])
            flags = 1 | (self.C << 8) | (Ce << 16)
            if is_grad_head:
                flags |= 2 if self.grad is not None else 4
            desc = [flags, self.eps.tensor.data_ptr(), self.x_in.tensor.data_ptr(), self.ent["t_loc"].data_ptr(),
                    d.sqrt_recip_alphas_cumprod.data_ptr(), d.sqrt_recip_alphas_cumprod_m1.data_ptr(),
                    d.sqrt_one_minus_alphas_cumprod.data_ptr(), tab.data_ptr()]
            fuse.tensor.copy_(torch.tensor(desc, dtype=torch.int64))
        gkey = "graph_fused" if self.fused else "graph"
        if use_graph and self.ent.get(gkey) is None:
            self.seek(1 if self.direction == "sample" else 0)
            self._launch_step()           # warm-up outside capture (lazy module loading, cudaFuncSetAttribute, ...)
            torch.cuda.synchronize(self.d.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch_step()
            self.ent[gkey] = g
        self.graph = self.ent.get(gkey)
        self.use_graph = use_graph

    def end(self):
        """Switch the fused update off again: the plan is shared with plain forward calls."""
        if getattr(self, "fuse_buf", None) is not None:
            self.fuse_buf.tensor.zero_()

    def seek(self, i: int):
        self.ent["counter"].fill_(int(i))

    def step(self):
        if self.use_graph:
            self.graph.replay()
        else:
            self._launch_step()
        if not self.in_graph_update:
            x = self.x_in.tensor
            e = self.eps.tensor[:, :self.C].contiguous()
            self.d._update(x, self.ent["t_loc"], e, self.grad.tensor if self.grad is not None else None, self.direction, out=x)


class DDIM:
    # replay each network step as one CUDA graph in the native fast path (PDAE_NO_GRAPH=1 disables, e.g. under ncu)
    use_cuda_graph = os.environ.get("PDAE_NO_GRAPH", "0") != "1"

    def __init__(self, betas, timestep_map, device):
        self.device = device
        self.timestep_map = timestep_map.to(self.device)
        self.timesteps = betas.shape[0] - 1
        # fp64 schedule algebra on the host, fp32 tables on the device -- exactly ddim.py:15-33
        acp = np.cumprod(1.0 - betas, axis=0)
        f32 = partial(torch.tensor, dtype=torch.float32, device=self.device)
        self.alphas_cumprod_prev = f32(np.append(1.0, acp[:-1]))
        self.alphas_cumprod_next = f32(np.append(acp[1:], 0.0))
        self.sqrt_one_minus_alphas_cumprod = f32(np.sqrt(1.0 - acp))
        self.sqrt_recip_alphas_cumprod = f32(np.sqrt(1.0 / acp))
        self.sqrt_recip_alphas_cumprod_m1 = f32(np.sqrt(1.0 / acp - 1.0))

    @staticmethod
    def extract_coef_at_t(schedule, t, x_shape):
        return torch.gather(schedule, -1, t).reshape([x_shape[0]] + [1] * (len(x_shape) - 1))

    def t_transform(self, t):
        return self.timestep_map[t]

    # ---- the fused update ---------------------------------------------------------------------------
    def _update(self, x_t, t, eps, grad, direction, out=None):
        if not x_t.is_cuda:
            raise _native.NativeError("DDIM: CUDA tensors required (no CPU fallback)")
        x_t, eps, grad = _f32(x_t), _f32(eps), _f32(grad)
        B = x_t.shape[0]
        t = _t64(t, B, "DDIM update")
        if eps.shape != x_t.shape or (grad is not None and grad.shape != x_t.shape):
            raise ValueError(f"DDIM update: x_t {tuple(x_t.shape)}, eps {tuple(eps.shape)} and grad must have one shape")
        out = torch.empty_like(x_t) if out is None else out
        tab = self.alphas_cumprod_prev if direction == "sample" else self.alphas_cumprod_next
        rc = _native.lib().pdae_ddim_step(_ptr(x_t), _ptr(eps), _ptr(grad), _ptr(t), _ptr(self.sqrt_recip_alphas_cumprod),
                                          _ptr(self.sqrt_recip_alphas_cumprod_m1), _ptr(self.sqrt_one_minus_alphas_cumprod),
                                          _ptr(tab), _ptr(out), B, x_t.numel() // B, _stream(x_t.device))
        _native.check(rc, "pdae_ddim_step")
        return out

    # ---- single steps (ddim.py:43-55, 66-79, 91-107, 123-138) -----------------------------------------
    def ddim_sample(self, denoise_fn, x_t, t, condition=None):
        return self._update(x_t, t, denoise_fn(x_t, self.t_transform(t), condition), None, "sample")

    def ddim_encode(self, denoise_fn, x_t, t, condition=None):
        return self._update(x_t, t, denoise_fn(x_t, self.t_transform(t), condition), None, "encode")

    def shift_ddim_sample(self, decoder, z, x_t, t, use_shift=True):
        eps, grad = decoder(x_t, self.t_transform(t), z)
        return self._update(x_t, t, eps, grad if use_shift else None, "sample")

    def shift_ddim_encode(self, decoder, z, x_t, t):
        eps, grad = decoder(x_t, self.t_transform(t), z)
        return self._update(x_t, t, eps, grad, "encode")

    # ---- loops (ddim.py:57-64, 81-88, 110-120, 140-147) -----------------------------------------------
    def _steps(self, direction):
        return reversed(range(1, self.timesteps + 1)) if direction == "sample" else range(0, self.timesteps)

    def _loop(self, net, x, cond, direction, shift, stop_step=0):
        from ..model.shift_unet import ShiftUNet
        from ..model.unet import UNet
        B = x.shape[0]
        fast = isinstance(net, (ShiftUNet, UNet)) and x.is_cuda and x.dim() == 4 and not torch.is_grad_enabled()
        ts = torch.arange(0, self.timesteps + 1, device=self.device, dtype=torch.int64)
        if not fast:
            img = x
            for i in self._steps(direction):
                t = ts[i].expand(B).contiguous()
                if shift:
                    eps, grad = net(img, self.t_transform(t), cond)
                    img = self._update(img, t, eps, grad if (direction == "encode" or (i - 1) >= stop_step) else None,
                                       direction)
                else:
                    img = self._update(img, t, net(img, self.t_transform(t), cond), None, direction)
            return img
        # fast path: drive the network's static plan buffers in place; a WHOLE step -- loop-index broadcast + timestep
        # map lookup, every network launch, the fused DDIM update writing x_t back into the network's input buffer -- is one
        # CUDA graph (device-side step counter), so a step costs one graph launch and no other host work
        H, W = x.shape[2], x.shape[3]
        C = x.shape[1]
        tail = None   # StepRunner of the epsilon-only plan used once the shift is switched off
        if isinstance(net, ShiftUNet):
            plan, (x_in, t_in, z_in, eps, grad) = net.plan_for(B, H, W)
            z_in.tensor.copy_(cond)
            main = _StepRunner(self, plan, x_in, t_in, eps, grad if shift else None, direction, C)
            if shift and direction == "sample" and stop_step > 0:
                # ddim.py:119: steps with (i-1) < stop_step ignore the shift -> replay only the frozen epsilon half there
                p2, (x2, t2, _, eps2, _) = net.plan_for(B, H, W, with_shift=False)
                tail = _StepRunner(self, p2, x2, t2, eps2, None, direction, C)
        else:
            plan, (x_in, t_in, c_in, eps) = net._get_plan(("unet", B, H, W), lambda P: net._build(P, B, H, W))
            if c_in is not None:
                c_in.tensor.copy_(cond)
            main = _StepRunner(self, plan, x_in, t_in, eps, None, direction, C)
        main.begin(self.use_cuda_graph)   # re-packs weights (forced: `.data` / raw-pointer updates bump no version), prologue
        if tail is not None:
            tail.begin(self.use_cuda_graph)
        try:
            x_in.tensor.copy_(x)
            cur = main
            steps = list(self._steps(direction))
            cur.seek(steps[0])
            for i in steps:
                use_shift = shift and (direction == "encode" or (i - 1) >= stop_step)
                if tail is not None and not use_shift and cur is not tail:
                    tail.x_in.tensor.copy_(cur.x_in.tensor)
                    cur = tail
                    cur.seek(i)
                cur.step()
            return cur.x_in.tensor.clone()
        finally:
            main.end()
            if tail is not None:
                tail.end()

    def ddim_sample_loop(self, denoise_fn, x_T, condition=None):
        return self._loop(denoise_fn, x_T, condition, "sample", shift=False)

    def ddim_encode_loop(self, denoise_fn, x_0, condition=None):
        return self._loop(denoise_fn, x_0, condition, "encode", shift=False)

    def shift_ddim_sample_loop(self, decoder, z, x_T, stop_percent=0.0):
        return self._loop(decoder, x_T, z, "sample", shift=True, stop_step=int(stop_percent * self.timesteps))

    def shift_ddim_encode_loop(self, decoder, z, x_0):
        return self._loop(decoder, x_0, z, "encode", shift=True)

    def shift_ddim_trajectory_interpolation(self, decoder, z_1, z_2, x_T, alpha):
        """ddim.py:149-174: two decoder calls per step, gradient = (1-alpha) g1 + alpha g2, epsilon from the first."""
        B = x_T.shape[0]
        x_t = x_T
        for i in reversed(range(1, self.timesteps + 1)):
            t = torch.full((B,), i, device=self.device, dtype=torch.long)
            eps, g1 = decoder(x_t, self.t_transform(t), z_1)
            _, g2 = decoder(x_t, self.t_transform(t), z_2)
            x_t = self._update(x_t, t, eps, (1.0 - alpha) * g1 + alpha * g2, "sample")
        return x_t

    def latent_ddim_sample(self, latent_denoise_fn, z_t, t):
        """ddim.py:178-198 -- the unclamped single step (no caller in the reference uses it: its loop goes through
        ddim_sample).  z_0 = A_t z_t - B_t eps;  z_prev = sqrt(abar_prev) z_0 + sqrt(1 - abar_prev) eps."""
        s = z_t.shape
        eps = latent_denoise_fn(z_t, self.t_transform(t))
        z0 = self.extract_coef_at_t(self.sqrt_recip_alphas_cumprod, t, s) * z_t - \
            self.extract_coef_at_t(self.sqrt_recip_alphas_cumprod_m1, t, s) * eps
        ap = self.extract_coef_at_t(self.alphas_cumprod_prev, t, s)
        return z0 * torch.sqrt(ap) + torch.sqrt(1.0 - ap) * eps

    def latent_ddim_sample_loop(self, latent_denoise_fn, z_T):
        """ddim.py:200-207 -- NB calls ddim_sample, i.e. WITH the clamp of the predicted z_0."""
        B = z_T.shape[0]
        z = z_T
        for i in reversed(range(1, self.timesteps + 1)):
            t = torch.full((B,), i, device=self.device, dtype=torch.long)
            z = self.ddim_sample(latent_denoise_fn, z, t)
        return z
