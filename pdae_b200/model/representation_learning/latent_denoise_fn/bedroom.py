from ...mlp_skip_net import MLPSkipNet

BEDROOMLatentDenoiseFn = MLPSkipNet  # reference: model/representation_learning/latent_denoise_fn/bedroom.py
