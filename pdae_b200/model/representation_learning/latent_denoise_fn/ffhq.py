from ...mlp_skip_net import MLPSkipNet

FFHQLatentDenoiseFn = MLPSkipNet  # reference: model/representation_learning/latent_denoise_fn/ffhq.py
