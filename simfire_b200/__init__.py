"""
simfire_b200 -- B200-native drop-in for SimFire's fire-spread hot path
(`RothermelFireManager.update`, simfire/game/managers/fire.py:616-719).

The compute lives in libsimfire_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/simfire_b200.h).  There is no CPU or PyTorch fallback: importing the package is
cheap, but constructing an engine without the library or without a GPU raises.
"""
from ._lib import SfbError  # noqa: F401
from .engine import FireEngine, rate_of_spread  # noqa: F401

__all__ = ["FireEngine", "SfbError", "rate_of_spread"]
__version__ = "0.1.0"
