"""Expose pdae_b200 under the reference's import names.

The reference's trainers/samplers do ``from model.unet import UNet``, ``import model.representation_learning.decoder
as decoder_module`` + ``getattr(decoder_module, cfg["model"])``, ``from diffusion.gaussian_diffusion import
GaussianDiffusion`` (trainer/train_representation_learning.py:10-17, sampler/autoencoding_eval.py:5-12).
``install()`` registers this package's modules in ``sys.modules`` under exactly those names, so the scripts pick up
the native implementation without edits.  Call it before importing any trainer/sampler module.
"""
import importlib
import sys

_ALIASES = {
    "model": "pdae_b200.model",
    "model.module": "pdae_b200.model.module",
    "model.unet": "pdae_b200.model.unet",
    "model.shift_unet": "pdae_b200.model.shift_unet",
    "model.mlp_skip_net": "pdae_b200.model.mlp_skip_net",
    "model.denoise_fn": "pdae_b200.model.denoise_fn",
    "model.representation_learning": "pdae_b200.model.representation_learning",
    "model.representation_learning.encoder": "pdae_b200.model.representation_learning.encoder",
    "model.representation_learning.decoder": "pdae_b200.model.representation_learning.decoder",
    "model.representation_learning.latent_denoise_fn": "pdae_b200.model.representation_learning.latent_denoise_fn",
    "diffusion": "pdae_b200.diffusion",
    "diffusion.gaussian_diffusion": "pdae_b200.diffusion.gaussian_diffusion",
    "diffusion.ddim": "pdae_b200.diffusion.ddim",
}


def install(force: bool = False) -> None:
    for alias, real in _ALIASES.items():
        if alias in sys.modules and not force and sys.modules[alias].__name__ != real:
            raise RuntimeError(f"'{alias}' is already imported from {sys.modules[alias].__file__}; call "
                               "pdae_b200.dropin.install() before importing the reference's scripts")
        sys.modules[alias] = importlib.import_module(real)
