"""Unconditional denoisers, looked up by name: ``getattr(model.denoise_fn, cfg['model'])``
(reference trainer/train_regular_diffusion.py:27)."""
from . import mnist

MNISTDenoiseFn = mnist.MNISTDenoiseFn

__all__ = ["MNISTDenoiseFn"]
