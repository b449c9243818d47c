"""Model configs of the parity tests (the definitions live in the package so bench.py does not import the test tree)."""
from pdae_b200.configs import *  # noqa: F401,F403
from pdae_b200.configs import _COMMON  # noqa: F401
