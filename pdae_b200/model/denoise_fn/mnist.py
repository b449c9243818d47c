from ..unet import UNet

MNISTDenoiseFn = UNet  # reference: model/denoise_fn/mnist.py
