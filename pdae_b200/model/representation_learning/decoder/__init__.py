from .celeba64 import CELEBA64Decoder
from .ffhq import FFHQDecoder
from .celebahq import CELEBAHQDecoder
from .bedroom import BEDROOMDecoder
from .horse import HORSEDecoder
