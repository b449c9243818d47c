"""Run a few ShiftUNet decoder steps (no CUDA graph) for Nsight Compute captures.
usage: ncu ... python scripts/ncu_step.py [workload] [batch] [precision] [n_steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pdae_b200
from bench import WORKLOADS
from pdae_b200.model.shift_unet import ShiftUNet
from pdae_b200.utils.synth import fill_module_, synth_normal

wl = sys.argv[1] if len(sys.argv) > 1 else "celeba64"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
cfg, size = WORKLOADS[wl][0], WORKLOADS[wl][1]
pdae_b200.set_default_precision(prec)
dev = torch.device("cuda")
dec = fill_module_(ShiftUNet(latent_dim=512, **cfg), seed=0).eval().to(dev)
x = synth_normal((B, 3, size, size), 1).to(dev)
z = synth_normal((B, 512), 2).to(dev)
t = torch.full((B,), 500, device=dev, dtype=torch.long)
with torch.no_grad():
    for _ in range(n):
        dec(x, t, z)
torch.cuda.synchronize()
plan, _ = dec.plan_for(B, size, size)
print("launches per step:", plan.n_launch, "arena MB:", plan.arena_bytes / 2 ** 20)
