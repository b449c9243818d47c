"""Latent DPM networks (MLPSkipNet subclasses), looked up by name (reference trainer/train_latent_diffusion.py)."""
from . import bedroom, celeba64, ffhq, horse

CELEBA64LatentDenoiseFn = celeba64.CELEBA64LatentDenoiseFn
FFHQLatentDenoiseFn = ffhq.FFHQLatentDenoiseFn
HORSELatentDenoiseFn = horse.HORSELatentDenoiseFn
BEDROOMLatentDenoiseFn = bedroom.BEDROOMLatentDenoiseFn

__all__ = ["CELEBA64LatentDenoiseFn", "FFHQLatentDenoiseFn", "HORSELatentDenoiseFn", "BEDROOMLatentDenoiseFn"]
