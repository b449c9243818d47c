from .common import Encoder128


class HORSEEncoder(Encoder128):
    """reference: model/representation_learning/encoder/horse.py"""
