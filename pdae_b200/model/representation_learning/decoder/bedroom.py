from ...shift_unet import ShiftUNet

BEDROOMDecoder = ShiftUNet  # reference: model/representation_learning/decoder/bedroom.py
