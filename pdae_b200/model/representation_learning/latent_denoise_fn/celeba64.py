from ...mlp_skip_net import MLPSkipNet

CELEBA64LatentDenoiseFn = MLPSkipNet  # reference: model/representation_learning/latent_denoise_fn/celeba64.py
