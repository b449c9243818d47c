"""Two whole PDAE training steps (forward, backward, fused Adam + EMA) for an `ncu` launch list: celeba64-proxy, B=32.
usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/ncu_train_step.py"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import WORKLOADS
from pdae_b200.diffusion.gaussian_diffusion import GaussianDiffusion
from pdae_b200.model.representation_learning.encoder import CELEBA64Encoder
from pdae_b200.model.shift_unet import ShiftUNet
from pdae_b200.optim import FusedAdamEMA
from pdae_b200.utils.synth import fill_module_, synth_images

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg, size = WORKLOADS["celeba64"][0], WORKLOADS["celeba64"][1]
dev = torch.device("cuda")
dec = fill_module_(ShiftUNet(latent_dim=512, **dict(cfg, dropout=0.1)), seed=0).to(dev)
enc = fill_module_(CELEBA64Encoder(latent_dim=512), seed=1).to(dev).train()
dec.freeze()
dec.set_train_mode()
dec.precision = enc.precision = "fp32"
ema_dec, ema_enc = copy.deepcopy(dec).requires_grad_(False), copy.deepcopy(enc).requires_grad_(False)
gd = GaussianDiffusion({"timesteps": 1000, "betas_type": "linear"}, dev)
groups = [list(enc.parameters()), list(dec.label_emb.parameters()), list(dec.shift_middle_block.parameters()),
          list(dec.shift_output_blocks.parameters()), list(dec.shift_out.parameters())]
opt = FusedAdamEMA([{"params": g} for g in groups], lr=1e-4, ema_decay=0.9999)
opt.attach_ema(enc, ema_enc)
opt.attach_ema(dec, ema_dec)
x0 = synth_images(B, 3, size, 3).to(dev)
for _ in range(2):
    loss = gd.representation_learning_train_one_batch(enc, dec, x0)["prediction_loss"]
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("loss", float(loss.detach()))
