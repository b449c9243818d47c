from .celeba64 import CELEBA64Encoder
from .ffhq import FFHQEncoder
from .celebahq import CELEBAHQEncoder
from .bedroom import BEDROOMEncoder
from .horse import HORSEEncoder
