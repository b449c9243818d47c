from ...shift_unet import ShiftUNet

FFHQDecoder = ShiftUNet  # reference: model/representation_learning/decoder/ffhq.py
