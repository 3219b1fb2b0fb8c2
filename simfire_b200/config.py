"""
`Config`: the subset of simfire/utils/config.py that feeds the fire-spread hot path, with
the same YAML schema, section names and field names (`config.area.screen_size`,
`config.simulation.update_rate`, `config.fire.fire_initial_position`, ...), so that a
SimFire YAML file or `Config(config_dict=...)` drives `simfire_b200.FireSimulation`.

Supported: functional topography (`flat`, `gaussian`), functional fuel (`chaparral`),
`simple` wind, static / random ignition -- what BASELINE config 1
(`configs/functional_config.yml`, flat) needs -- plus `Config.from_arrays` for callers that
already hold terrain arrays (operational LANDFIRE data, Perlin or CFD wind produced
elsewhere).  The data-ingest layers themselves (LANDFIRE download, BurnMD, simplex noise,
CFD) are outside the hot-path scope and raise
`ConfigError` here.
"""
from __future__ import annotations

import re
import warnings
from dataclasses import dataclass, field
from datetime import timedelta
from math import exp
from pathlib import Path
from typing import Any, Dict, Optional, Tuple, Union

import numpy as np
import yaml

from .enums import FuelConstants
from .parameters import Fuel
from .workloads import MPH_TO_FTPM


class ConfigError(Exception):
    """simfire/utils/config.py (ConfigError)"""


_UNITS = {"s": "seconds", "m": "minutes", "h": "hours", "d": "days", "w": "weeks"}


def str_to_minutes(string: str) -> int:
    """'1d 2h 3m' -> minutes (simfire/utils/units.py:62-86)."""
    return int(round(timedelta(**{
        _UNITS.get(m.group("unit").lower(), "minutes"): float(m.group("val"))
        for m in re.finditer(r"(?P<val>\d+(\.\d+)?)(?P<unit>[smhdw]?)", string, flags=re.I)
    }).total_seconds() / 60))  # fmt: skip


def mph_to_ftpm(mph: float) -> float:
    return mph * MPH_TO_FTPM


def chaparral(seed: Optional[int] = None) -> Fuel:
    """simfire/utils/terrain.py:93-114: every field re-seeds NumPy's global generator, so one
    uniform draw u sets all four values."""
    def draw(lo, hi):
        np.random.seed(seed)
        return np.random.uniform(lo, hi)

    return Fuel(w_0=draw(FuelConstants.W_0_MIN, FuelConstants.W_0_MAX),
                delta=draw(FuelConstants.DELTA_MIN, FuelConstants.DELTA_MAX),
                M_x=draw(FuelConstants.M_X_MIN, FuelConstants.M_X_MAX),
                sigma=draw(FuelConstants.SIGMA_MIN, FuelConstants.SIGMA_MAX))  # fmt: skip


@dataclass
class AreaConfig:
    screen_size: Tuple[int, int]
    pixel_scale: float

    def __post_init__(self):
        self.screen_size = (int(self.screen_size[0]), int(self.screen_size[1]))
        self.pixel_scale = float(self.pixel_scale)


@dataclass
class DisplayConfig:
    fire_size: int = 2
    control_line_size: int = 2
    agent_size: int = 4
    rescale_factor: Optional[int] = None


@dataclass
class SimulationConfig:
    update_rate: float
    runtime: Union[str, int]
    headless: bool = True
    draw_spread_graph: bool = False
    record: bool = False
    save_data: bool = False
    data_type: str = "npy"
    sf_home: str = "~/.simfire"

    def __post_init__(self):
        self.update_rate = float(self.update_rate)
        if isinstance(self.runtime, str):  # config.py:_load_simulation
            self.runtime = str_to_minutes(self.runtime)
        self.sf_home = Path(self.sf_home).expanduser()


@dataclass
class MitigationConfig:
    ros_attenuation: bool = True


@dataclass
class _ArrayLayer:
    """Stand-in for a simfire data layer: only `.data` is read downstream."""

    data: np.ndarray
    name: str = "array"


@dataclass
class FunctionSpec:
    """Name + keyword arguments of a functional layer generator (the reference keeps the
    callable too; `.name` and `.kwargs` are what FireSimulation.get_seeds reads)."""

    name: str
    kwargs: Dict[str, Any]


@dataclass
class TerrainConfig:
    topography_type: str
    topography_layer: _ArrayLayer  # .data: (H, W, 1) elevations in ft
    fuel_type: str
    fuel_layer: _ArrayLayer  # .data: (H, W, 1) object array of Fuel
    topography_function: Optional[FunctionSpec] = None
    fuel_function: Optional[FunctionSpec] = None


@dataclass
class FireConfig:
    fire_initial_position: Tuple[int, int]
    diagonal_spread: bool
    max_fire_duration: int
    seed: Optional[int] = None


@dataclass
class EnvironmentConfig:
    moisture: float


@dataclass
class WindConfig:
    speed: np.ndarray  # ft/min, (H, W) float64 (config.py:943-944)
    direction: np.ndarray  # degrees clockwise from north
    speed_function: Optional[FunctionSpec] = None
    direction_function: Optional[FunctionSpec] = None


class Config:
    def __init__(self, path: Optional[Union[str, Path]] = None, config_dict: Optional[Dict[str, Any]] = None,
                 *, _sections: Optional[Dict[str, Any]] = None) -> None:  # fmt: skip
        """Either a YAML `path` or an equivalent `config_dict` (config.py:208-270)."""
        if _sections is not None:
            self.yaml_data = {}
            self.path = None
            for k, v in _sections.items():
                setattr(self, k, v)
            return
        if path is not None:
            self.path = Path(path)
            with open(self.path) as f:
                self.yaml_data = yaml.safe_load(f)
        elif config_dict is not None:
            self.path = None
            self.yaml_data = config_dict
        else:
            raise ConfigError("Either `path` or `config_dict` must be supplied to Config")
        y = self.yaml_data
        self.area = AreaConfig(**y["area"])
        self.display = DisplayConfig(**y.get("display", {}))
        self.simulation = SimulationConfig(**y["simulation"])
        self.mitigation = MitigationConfig(**y.get("mitigation", {}))
        self.terrain = self._load_terrain()
        self.fire = self._load_fire()
        self.environment = EnvironmentConfig(**y["environment"])
        self.wind = self._load_wind()

    # -- sections -------------------------------------------------------------------------
    def _load_terrain(self) -> TerrainConfig:
        H, W = self.area.screen_size
        t = self.yaml_data["terrain"]
        topo = t["topography"]
        if topo["type"] != "functional":
            raise ConfigError(f"topography type {topo['type']!r}: only 'functional' terrain is built here; pass "
                              "arrays with Config.from_arrays for operational / historical data")  # fmt: skip
        fn = topo["functional"]["function"]
        xx, yy = np.meshgrid(np.arange(W), np.arange(H))
        if fn == "flat":
            elev = np.zeros((H, W), dtype=np.int64)  # the reference's flat() returns the int 0
            kwargs = {}
        elif fn == "gaussian":
            kwargs = dict(topo["functional"]["gaussian"])
            g = np.vectorize(lambda x, y: kwargs["amplitude"] * exp(-(
                ((x - kwargs["mu_x"]) ** 2 / (4 * kwargs["sigma_x"] ** 2))
                + ((y - kwargs["mu_y"]) ** 2 / (4 * kwargs["sigma_y"] ** 2)))))  # elevation_functions.py:33-70
            elev = g(xx, yy)
        else:
            raise ConfigError(f"topography function {fn!r} needs the `noise` package (simplex noise); build the "
                              "elevation array elsewhere and use Config.from_arrays")  # fmt: skip
        fuel = t["fuel"]
        if fuel["type"] != "functional":
            raise ConfigError(f"fuel type {fuel['type']!r}: only 'functional' fuel is built here")
        ffn = fuel["functional"]["function"]
        if ffn != "chaparral":
            raise ConfigError(f"fuel function {ffn!r} is not supported")
        fkw = dict(fuel["functional"]["chaparral"])
        the_fuel = chaparral(**fkw)
        fuels = np.empty((H, W), dtype=object)
        fuels.fill(the_fuel)
        return TerrainConfig("functional", _ArrayLayer(elev[..., None], fn), "functional",
                             _ArrayLayer(fuels[..., None], ffn), FunctionSpec(fn, kwargs),
                             FunctionSpec(ffn, fkw))  # fmt: skip

    def _load_fire(self, pos: Optional[Tuple[int, int]] = None) -> FireConfig:
        """config.py:775-827"""
        f = self.yaml_data["fire"]
        max_dur, diag = int(f["max_fire_duration"]), bool(f["diagonal_spread"])
        kind = f["fire_initial_position"]["type"]
        if kind == "static":
            if pos is None:
                pos = f["fire_initial_position"]["static"]["position"]
                if isinstance(pos, str):
                    pos = pos[1:-1].split(",")
                if len(pos) > 2:
                    raise ConfigError("`fire_initial_position` should only be a Tuple of length 2")
            return FireConfig((int(pos[0]), int(pos[1])), diag, max_dur)
        if kind == "random":
            if pos is not None:
                warnings.warn("`pos` is specified, but the initialization type is `random`. Ignoring `pos`.")
            seed = f["fire_initial_position"]["random"]["seed"]
            rng = np.random.default_rng(seed)  # config.py:808-812: x first, then y
            H, W = self.yaml_data["area"]["screen_size"]
            pos_x = int(rng.integers(W, dtype=int))
            pos_y = int(rng.integers(H, dtype=int))
            return FireConfig((pos_x, pos_y), diag, max_dur, seed)
        raise ConfigError(f"The specified fire initial position type ({kind}) is not supported")

    def _load_wind(self) -> WindConfig:
        w = self.yaml_data["wind"]
        if w["function"] != "simple":
            raise ConfigError(f"wind function {w['function']!r} is produced by the reference's wind generators "
                              "(simplex noise / CFD); pass the arrays with Config.from_arrays")  # fmt: skip
        shape = self.area.screen_size
        speed = np.full(shape, mph_to_ftpm(w["simple"]["speed"])).astype(np.float64)
        direction = np.full(shape, w["simple"]["direction"]).astype(np.float64)
        return WindConfig(speed, direction)  # `simple` wind carries no function spec (config.py:864-865)

    # -- array entry point ------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, *, fuels: np.ndarray, elevations: np.ndarray, wind_speed, wind_direction,
                    pixel_scale: float, fire_initial_position: Tuple[int, int], update_rate: float = 1.0,
                    runtime: Union[str, int] = "24h", max_fire_duration: int = 4, diagonal_spread: bool = True,
                    moisture: float = 0.03, ros_attenuation: bool = True) -> "Config":  # fmt: skip
        """
        Build a Config from terrain arrays.  `fuels`: (H, W) object array of `Fuel` or an
        (H, W, 4) float array (w_0, delta, M_x, sigma); `elevations`: (H, W) ft;
        `wind_speed` (ft/min) / `wind_direction` (deg): scalars or (H, W) arrays.
        """
        fuels = np.asarray(fuels)
        H, W = fuels.shape[:2]
        speed = np.broadcast_to(np.asarray(wind_speed, dtype=np.float64), (H, W)).copy()
        direction = np.broadcast_to(np.asarray(wind_direction, dtype=np.float64), (H, W)).copy()
        fdata = fuels[..., None] if fuels.dtype == object else fuels
        sections = dict(
            area=AreaConfig((H, W), pixel_scale), display=DisplayConfig(),
            simulation=SimulationConfig(update_rate, runtime), mitigation=MitigationConfig(ros_attenuation),
            terrain=TerrainConfig("arrays", _ArrayLayer(np.asarray(elevations).reshape(H, W, 1)), "arrays",
                                  _ArrayLayer(fdata)),
            fire=FireConfig(tuple(int(v) for v in fire_initial_position), diagonal_spread, max_fire_duration),
            environment=EnvironmentConfig(moisture), wind=WindConfig(speed, direction),
        )  # fmt: skip
        cfg = cls(_sections=sections)
        # enough YAML for reset_fire() to work the way it does on a file-backed Config
        cfg.yaml_data = {
            "area": {"screen_size": [H, W], "pixel_scale": pixel_scale},
            "fire": {"fire_initial_position": {"type": "static", "static": {"position": tuple(sections["fire"].fire_initial_position)},
                                               "random": {"seed": None}},
                     "max_fire_duration": max_fire_duration, "diagonal_spread": diagonal_spread},
        }  # fmt: skip
        return cfg

    # -- mutation between episodes (config.py:975-1133) -------------------------------------
    def reset_terrain(self, topography_seed: Optional[int] = None, topography_type: Optional[str] = None,
                      fuel_seed: Optional[int] = None, fuel_type: Optional[str] = None,
                      location: Optional[Tuple[float, float]] = None) -> None:  # fmt: skip
        """Rewrite the functional seeds / layer types in the YAML data and rebuild the terrain
        (config.py:975-1046).  Array-backed configs have nothing to regenerate from."""
        if "terrain" not in self.yaml_data:
            raise ConfigError("reset_terrain: this Config was built from arrays; build a new one with from_arrays")
        if location is not None:
            raise ConfigError("reset_terrain(location=...): operational (LANDFIRE) layers are not built here")
        t = self.yaml_data["terrain"]
        if topography_seed is not None and self.terrain.topography_function is not None:
            # KeyError for 'flat', which has no block in the YAML -- as in the reference (config.py:1010)
            t["topography"]["functional"][self.terrain.topography_function.name]["seed"] = topography_seed
        if fuel_seed is not None and self.terrain.fuel_function is not None:
            t["fuel"]["functional"][self.terrain.fuel_function.name]["seed"] = fuel_seed
        if topography_type is not None:
            t["topography"]["type"] = topography_type
        if fuel_type is not None:
            t["fuel"]["type"] = fuel_type
        self.area = AreaConfig(**self.yaml_data["area"])
        self.terrain = self._load_terrain()

    def reset_wind(self, speed_seed: Optional[int] = None, direction_seed: Optional[int] = None) -> None:
        """config.py:1048-1086.  `simple` wind has no seed (speed_function is None), so the seeds
        are ignored exactly as the reference ignores them and the same field is reloaded."""
        if "wind" not in self.yaml_data:
            raise ConfigError("reset_wind: this Config was built from arrays; build a new one with from_arrays")
        self.wind = self._load_wind()

    def reset_fire(self, seed: Optional[int] = None, pos: Optional[Tuple[int, int]] = None) -> None:
        """config.py:1088-1133: `seed` re-draws a `random` start, `pos` moves a `static` one; the
        other combination is ignored with a warning, both or neither is a ValueError."""
        kind = self.yaml_data["fire"]["fire_initial_position"]["type"]
        if seed is None and pos is None:
            raise ValueError("Both `seed` and `pos` cannot be None")
        if seed is not None and pos is not None:
            raise ValueError("Both `seed` and `pos` cannot be specified together")
        key, value = ("seed", seed) if seed is not None else ("position", pos)
        section = self.yaml_data["fire"]["fire_initial_position"][kind]
        if (kind == "random") != (key == "seed"):
            # the reference stores the value, then _load_fire ignores it for this type
            warnings.warn(f"Trying to set a {key} for fire initial position type ({kind}), which does not "
                          f"support the use of a {key}. The {key} value will be ignored.")  # fmt: skip
            return
        section[key] = value
        self.fire = self._load_fire(pos=pos)
