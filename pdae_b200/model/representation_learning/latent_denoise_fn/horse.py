from ...mlp_skip_net import MLPSkipNet

HORSELatentDenoiseFn = MLPSkipNet  # reference: model/representation_learning/latent_denoise_fn/horse.py
