from .common import Encoder64


class CELEBA64Encoder(Encoder64):
    """reference: model/representation_learning/encoder/celeba64.py:4-37"""
