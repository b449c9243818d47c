from .mnist import MNISTDenoiseFn
