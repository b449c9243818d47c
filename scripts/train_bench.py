"""Training-step throughput: representation_learning_train_one_batch + backward + fused
Adam/EMA step on the celeba64-proxy decoder + encoder, fp32 CUDA-core arithmetic (the training path's mode).
usage: python scripts/train_bench.py [batch] [steps] [--cpu-oracle N]   (N = batch of the CPU oracle comparison, 0 = skip)"""
import copy
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import WORKLOADS
from pdae_b200.diffusion.gaussian_diffusion import GaussianDiffusion
from pdae_b200.model.representation_learning.encoder import CELEBA64Encoder
from pdae_b200.model.shift_unet import ShiftUNet
from pdae_b200.optim import FusedAdamEMA
from pdae_b200.utils.synth import fill_module_, synth_images

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cpu_b = int(sys.argv[sys.argv.index("--cpu-oracle") + 1]) if "--cpu-oracle" in sys.argv else 0
cfg, size = WORKLOADS["celeba64"][0], WORKLOADS["celeba64"][1]
dev = torch.device("cuda")
dec = fill_module_(ShiftUNet(latent_dim=512, **dict(cfg, dropout=0.1)), seed=0).to(dev)
enc = fill_module_(CELEBA64Encoder(latent_dim=512), seed=1).to(dev).train()
dec.freeze()
dec.set_train_mode()
dec.precision = enc.precision = "fp32"
ema_dec, ema_enc = copy.deepcopy(dec).requires_grad_(False), copy.deepcopy(enc).requires_grad_(False)
gd = GaussianDiffusion({"timesteps": 1000, "betas_type": "linear"}, dev)
opt = FusedAdamEMA([{"params": enc.parameters()}, {"params": dec.label_emb.parameters()},
                    {"params": dec.shift_middle_block.parameters()}, {"params": dec.shift_output_blocks.parameters()},
                    {"params": dec.shift_out.parameters()}], lr=1e-4, ema_decay=0.9999)
opt.attach_ema(enc, ema_enc)
opt.attach_ema(dec, ema_dec)
x0 = synth_images(B, 3, size, 3).to(dev)


def step():
    loss = gd.representation_learning_train_one_batch(enc, dec, x0)["prediction_loss"]
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"train_bench celeba64-proxy B={B}: {ms:.1f} ms/step  {B / ms * 1e3:.1f} img/s  (fp32 CUDA-core fwd+bwd + fused Adam/EMA; "
      f"loss {float(loss.detach()):.4f})")
e0.record()
for _ in range(20):
    for p in enc.parameters():
        p.grad = torch.zeros_like(p) if p.grad is None else p.grad
    opt.step()
e1.record()
torch.cuda.synchronize()
print(f"  fused Adam+EMA step alone (encoder group only has grads): {e0.elapsed_time(e1) / 20:.3f} ms")
if cpu_b:
    from oracle import pdae_oracle as O
    from pdae_b200.utils.host import host_cores
    torch.set_num_threads(host_cores())
    dsd = {k: v.detach().float().cpu().requires_grad_(k.startswith(("label_emb", "shift_"))) for k, v in dec.state_dict().items()}
    esd = {k: v.detach().float().cpu().requires_grad_(True) for k, v in enc.state_dict().items()}
    D = O.DiffusionOracle({"timesteps": 1000, "betas_type": "linear"})
    xc = x0[:cpu_b].cpu()
    t = torch.randint(0, 1000, (cpu_b,))
    noise = torch.randn_like(xc)
    t0 = time.time()
    l = D.representation_learning_loss(lambda x: O.encoder_forward(esd, "celeba64", x),
                                       lambda x, tt, z: O.shiftunet_forward(dsd, cfg, x, tt, z), xc, t, noise)
    l.backward()
    dt = time.time() - t0
    print(f"  CPU oracle (torch autograd, {host_cores()} threads) B={cpu_b}: {dt:.1f} s/step  {cpu_b / dt:.2f} img/s")
