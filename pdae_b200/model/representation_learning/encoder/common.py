"""Semantic encoders x0 -> z (reference: model/representation_learning/encoder/*.py).

Five near-identical stride-2 conv stacks; CelebA64 is the 4-stage / 64 px variant (celeba64.py:10-37), the
others the 5-stage / 128 px variant (ffhq.py:10-41 == celebahq / bedroom / horse).  The nn.Sequential index
numbering of the reference (``encoder.{idx}``) is reproduced with explicit slots.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

from ....engine import Plan, RESAMPLE_NONE
from ...module import AttentionBlock, PlannedModule, Slots, Src, normalization


class ConvStackEncoder(PlannedModule):
    widths: List[int] = []
    attn_after = 0      # index into widths after which the AttentionBlock sits
    image_size = 0

    def __init__(self, **kwargs):
        super().__init__()
        self.latent_dim = kwargs["latent_dim"]
        w = self.widths
        slots = {0: nn.Conv2d(3, w[0], (3, 3), (2, 2), 1)}
        self._order = [("conv", 0)]
        idx, cin = 1, w[0]
        for j in range(1, len(w)):
            slots[idx] = normalization(cin)
            slots[idx + 2] = nn.Conv2d(cin, w[j], (3, 3), (2, 2), 1)
            self._order += [("gn", idx), ("conv", idx + 2)]
            idx += 3
            cin = w[j]
            if j == self.attn_after:
                slots[idx] = AttentionBlock(cin, 4, -1, False)
                self._order.append(("attn", idx))
                idx += 1
        slots[idx] = normalization(cin)
        self._order.append(("gn", idx))
        idx += 3  # GroupNorm, SiLU, View
        slots[idx] = nn.Linear(cin * 16, self.latent_dim)
        self._order.append(("linear", idx))
        self.encoder = Slots(slots)

    def _build(self, P: Plan, B: int, H: int, W: int):
        x_in = P.new((B, 3, H, W), torch.float32, "x_nchw")
        x_in.keep = True
        h = None
        C = 3
        pending_ab = None  # (ab) of a GN whose SiLU output feeds the next conv
        z = None
        for kind, idx in self._order:
            m = self.encoder[idx]
            if kind == "conv":
                Co = m.weight.shape[0]
                Ho, Wo = H // 2, W // 2
                out = P.new((B, Ho, Wo, Co), torch.float32, "enc_h")
                if h is None:
                    P.conv(x_in, m.weight, m.bias, out, B=B, H=H, W=W, Cin=3, Cout=Co, k=3, stride=2, pad=1, in_nchw=True)
                else:
                    act, _ = P.gn_apply(h.b1, C, None, 0, pending_ab, silu=True, resample=RESAMPLE_NONE, B=B, H=H, W=W,
                                        act_dtype=torch.float32)
                    P.conv(act, m.weight, m.bias, out, B=B, H=H, W=W, Cin=C, Cout=Co, k=3, stride=2, pad=1)
                h, C, H, W = Src(out, Co, B, Ho, Wo), Co, Ho, Wo
            elif kind == "gn":
                pending_ab = P.gn_coef(h.b1, C, None, 0, m.weight, m.bias, B=B, HW=H * W)
            elif kind == "attn":
                h = m.emit(P, h)
            else:  # final GN+SiLU, View(-1, C*4*4) in NCHW order, Linear
                act, _ = P.gn_apply(h.b1, C, None, 0, pending_ab, silu=True, resample=RESAMPLE_NONE, B=B, H=H, W=W,
                                    act_dtype=torch.float32)
                # the reference flattens NCHW (c, y, x); our activation is NHWC (y, x, c): permute the weight instead
                wt = m.weight
                HW = H * W
                wp = P.pack((id(wt), "enc_fc"), [wt],
                            lambda: wt.detach().reshape(-1, C, HW).permute(2, 1, 0).reshape(HW * C, -1).float())
                z = P.new((B, self.latent_dim), torch.float32, "z")
                z.keep = True
                P.linear_packed(act, wp, P.param(m.bias), z, B=B, Cin=HW * C, Cout=self.latent_dim)
        return x_in, z

    def forward(self, x):
        """x [N,3,S,S] fp32 -> z [N, latent_dim]."""
        B, C, H, W = x.shape
        assert C == 3
        if (H, W) != (self.image_size, self.image_size):
            raise ValueError(f"{type(self).__name__} is hard-wired to {self.image_size}x{self.image_size} inputs "
                             f"(View(-1, C*4*4) in the reference), got {H}x{W}")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from ....train import encoder_train_forward
            return encoder_train_forward(self, x.contiguous())
        plan, (x_in, z) = self._get_plan(("enc", B, H, W), lambda P: self._build(P, B, H, W))
        x_in.tensor.copy_(x)
        plan.run()
        return z.tensor.clone()


class Encoder64(ConvStackEncoder):
    widths = [64, 128, 128, 128]
    attn_after = 1
    image_size = 64


class Encoder128(ConvStackEncoder):
    widths = [64, 128, 256, 256, 256]
    attn_after = 2
    image_size = 128
