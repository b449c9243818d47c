from .common import Encoder128


class BEDROOMEncoder(Encoder128):
    """reference: model/representation_learning/encoder/bedroom.py"""
