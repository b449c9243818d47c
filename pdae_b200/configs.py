"""Model configs used by bench.py, the driver entry points and the tests.

The decoder hyper-parameters of the reference live in un-shipped ``pre-trained-dpms/*/config.yml`` files (SURVEY.md
D4), so the *_PROXY configs below are assumptions that follow the guided-diffusion / Diff-AE lineage (SURVEY.md
section 8d); MNIST and the latent MLP are the reference's own (config/mnist_regular.yml:14-28, config/ffhq_latent.yml:16-23).
"""
_COMMON = dict(dims=2, input_channel=3, num_residual_blocks_of_a_block=2, num_heads=1, head_channel=-1,
               use_new_attention_order=False, dropout=0.0, learn_sigma=False)
MNIST = dict(model="MNISTDenoiseFn", dims=2, input_channel=1, base_channel=64, channel_multiplier=[1, 2, 2, 4],
             num_residual_blocks_of_a_block=2, dropout=0.0, attention_resolutions=[], use_new_attention_order=False,
             num_heads=1, head_channel=-1)
CELEBA64_PROXY = dict(_COMMON, base_channel=64, channel_multiplier=[1, 2, 4, 8], attention_resolutions=[4])
FFHQ128_PROXY = dict(_COMMON, base_channel=128, channel_multiplier=[1, 1, 2, 3, 4], attention_resolutions=[8])
FFHQ256_PROXY = dict(_COMMON, base_channel=128, channel_multiplier=[1, 1, 2, 2, 4, 4], attention_resolutions=[16])
FFHQ_LATENT = dict(model="FFHQLatentDenoiseFn", input_channel=512, model_channel=2048, num_layers=10, time_emb_channel=64,
                   use_norm=True, dropout=0.0)
DIFFUSION = {"timesteps": 1000, "betas_type": "linear"}
