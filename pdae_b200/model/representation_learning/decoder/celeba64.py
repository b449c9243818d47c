from ...shift_unet import ShiftUNet

CELEBA64Decoder = ShiftUNet  # reference: model/representation_learning/decoder/celeba64.py
