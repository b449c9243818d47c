from .common import Encoder128


class CELEBAHQEncoder(Encoder128):
    """reference: model/representation_learning/encoder/celebahq.py"""
