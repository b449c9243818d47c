"""Conditional decoders (ShiftUNet subclasses), looked up by name like the encoders
(reference trainer/train_representation_learning.py:28-38)."""
from . import bedroom, celeba64, celebahq, ffhq, horse

CELEBA64Decoder = celeba64.CELEBA64Decoder
FFHQDecoder = ffhq.FFHQDecoder
CELEBAHQDecoder = celebahq.CELEBAHQDecoder
BEDROOMDecoder = bedroom.BEDROOMDecoder
HORSEDecoder = horse.HORSEDecoder

__all__ = ["CELEBA64Decoder", "FFHQDecoder", "CELEBAHQDecoder", "BEDROOMDecoder", "HORSEDecoder"]
