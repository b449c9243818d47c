from ...shift_unet import ShiftUNet

HORSEDecoder = ShiftUNet  # reference: model/representation_learning/decoder/horse.py
