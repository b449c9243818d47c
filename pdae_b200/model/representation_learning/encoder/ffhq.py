from .common import Encoder128


class FFHQEncoder(Encoder128):
    """reference: model/representation_learning/encoder/ffhq.py"""
