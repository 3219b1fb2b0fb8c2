"""Dataclasses with the field names of simfire/world/parameters.py, so code that builds a
`RothermelFireManager` from `FuelParticle`, `Fuel` and `Environment` objects keeps working."""
from dataclasses import dataclass
from typing import Sequence, Union

import numpy as np


@dataclass
class FuelParticle:
    """simfire/world/parameters.py:8-27"""

    h: float = 8000.0  # low heat content (BTU/lb)
    S_T: float = 0.0555  # total mineral content
    S_e: float = 0.01  # effective mineral content
    p_p: float = 32.0  # oven-dry particle density (lb/ft^3)


@dataclass
class Fuel:
    """simfire/world/parameters.py:31-50"""

    w_0: float  # oven-dry fuel load (lb/ft^2)
    delta: float  # fuel bed depth (ft)
    M_x: float  # dead fuel moisture of extinction
    sigma: float  # surface-area-to-volume ratio (ft^2/ft^3)


@dataclass
class Environment:
    """simfire/world/parameters.py:53-76"""

    M_f: float  # fuel moisture
    U: Union[float, Sequence[Sequence[float]], np.ndarray]  # wind speed (ft/min)
    U_dir: Union[float, Sequence[Sequence[float]], np.ndarray]  # wind direction (degrees cw from north)
