"""Value encodings shared with SimFire (simfire/enums.py): the integers that appear in
`fire_map`, the status `update()` returns and the control-line attenuation constants."""
from enum import IntEnum


class BurnStatus(IntEnum):
    """simfire/enums.py:52-69"""

    UNBURNED = 0
    BURNING = 1
    BURNED = 2
    FIRELINE = 3
    SCRATCHLINE = 4
    WETLINE = 5


class GameStatus(IntEnum):
    """simfire/enums.py:106-115"""

    QUIT = 0
    RUNNING = 1


class RoSAttenuation(IntEnum):
    """Rate-of-spread attenuation of each control line in ft/min (simfire/enums.py:72-85).
    The device constants live in csrc/sfb_kernels.cuh (`line_attenuation`)."""

    FIRELINE = 980
    SCRATCHLINE = 490
    WETLINE = 245


class FuelConstants:
    """Observation-space bounds (simfire/enums.py:119-138)."""

    W_0_MIN = 0.0
    W_0_MAX = 1.0
    DELTA_MIN = 0.2
    DELTA_MAX = 6.0
    M_X_MIN = 0.12
    M_X_MAX = 1.0
    SIGMA_MIN = 1
    SIGMA_MAX = 3500


class ElevationConstants:
    """simfire/enums.py (ElevationConstants): Death Valley .. the treeline, in ft."""

    MIN_ELEVATION = -282
    MAX_ELEVATION = 11_000


class WindConstants:
    """simfire/enums.py (WindConstants), mph."""

    MIN_SPEED = 0
    MAX_SPEED = 250
