"""Semantic encoders, looked up by name: ``getattr(model.representation_learning.encoder, cfg['model'])``
(reference trainer/train_representation_learning.py:28-38).  One class per dataset config of the reference."""
from . import bedroom, celeba64, celebahq, ffhq, horse

CELEBA64Encoder = celeba64.CELEBA64Encoder
FFHQEncoder = ffhq.FFHQEncoder
CELEBAHQEncoder = celebahq.CELEBAHQEncoder
BEDROOMEncoder = bedroom.BEDROOMEncoder
HORSEEncoder = horse.HORSEEncoder

__all__ = ["CELEBA64Encoder", "FFHQEncoder", "CELEBAHQEncoder", "BEDROOMEncoder", "HORSEEncoder"]
