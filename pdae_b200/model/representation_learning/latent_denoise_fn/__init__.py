from .celeba64 import CELEBA64LatentDenoiseFn
from .ffhq import FFHQLatentDenoiseFn
from .horse import HORSELatentDenoiseFn
from .bedroom import BEDROOMLatentDenoiseFn
