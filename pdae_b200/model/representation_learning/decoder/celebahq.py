from ...shift_unet import ShiftUNet

CELEBAHQDecoder = ShiftUNet  # reference: model/representation_learning/decoder/celebahq.py
