"""
Drop-in for `simfire.game.managers.fire.RothermelFireManager` (fire.py:287-719) whose
`update()` runs on the B200 through libsimfire_b200.so.

Same constructor, same `update(fire_map) -> (fire_map, GameStatus)`, same attributes a
`FireSimulation` reads (`elapsed_time`, `sprites`, `burn_amounts`, `rate_of_spread`, `U`,
`U_dir`, `slope_mag`, `slope_dir`).  Differences, all outside the hot path:
  * no pygame sprites: `sprites` is a list of (x, y) tuples of the burning cells, built on
    demand from the device state (the reference uses it for rendering only,
    simulation.py:291, :534);
  * `fs_graph` is not maintained per step: with `keep_spread_graph=True` it is rebuilt on demand
    from the device's ignition-step plane (simfire_b200/graph.py).
"""
from __future__ import annotations

import collections.abc
from typing import Any, Optional, Tuple

import numpy as np

from .engine import FireEngine
from .enums import BurnStatus, GameStatus
from .parameters import Environment, FuelParticle
from .workloads import compute_slopes


def fuel_planes(fuels: np.ndarray):
    """(H, W) object array of `Fuel`-like objects -> w_0, delta, M_x, sigma float64 planes
    (what `_accrue_sprites` reads per pair with dataclasses.astuple, fire.py:481-483)."""
    fuels = np.asarray(fuels)
    if fuels.dtype != object:  # already (H, W, 4) numbers
        if fuels.ndim != 3 or fuels.shape[-1] != 4:
            raise ValueError("fuels must be an (H, W) array of Fuel objects or an (H, W, 4) array")
        return tuple(np.ascontiguousarray(fuels[..., i], dtype=np.float64) for i in range(4))
    H, W = fuels.shape
    out = np.empty((4, H, W), dtype=np.float64)
    cache = {}
    for y in range(H):
        for x in range(W):
            f = fuels[y, x]
            v = cache.get(id(f))
            if v is None:
                v = cache[id(f)] = (f.w_0, f.delta, f.M_x, f.sigma)
            out[:, y, x] = v
    return out[0], out[1], out[2], out[3]


class RothermelFireManager:
    def __init__(
        self,
        init_pos: Tuple[int, int],
        fire_size: int,
        max_fire_duration: int,
        pixel_scale: float,
        update_rate: float,
        fuel_particle: FuelParticle,
        terrain: Any,
        environment: Environment,
        max_time: Optional[int] = None,
        attenuate_line_ros: bool = True,
        headless: bool = False,
        diagonal_spread: bool = True,
        *,
        device: int = 0,
        keep_rate_of_spread: bool = False,
        keep_spread_graph: bool = False,
    ) -> None:
        """
        Arguments as in fire.py:293-366.  `terrain` needs `.fuels` ((H, W) array of objects
        with w_0 / delta / M_x / sigma), `.elevations` ((H, W) floats) and `.screen_size`.
        `keep_rate_of_spread=True` materialises the dense `rate_of_spread` attribute every
        step (fire.py:704-708) at the cost of an extra pass over the grid.
        """
        self.init_pos = tuple(int(v) for v in init_pos)
        self.fire_size = fire_size
        self.max_fire_duration = max_fire_duration
        self.attenuate_line_ros = attenuate_line_ros
        self.headless = headless
        self.diagonal_spread = diagonal_spread
        self.pixel_scale = pixel_scale
        self.update_rate = update_rate
        self.max_time = max_time
        self.fuel_particle = fuel_particle
        self.terrain = terrain
        self.environment = environment
        self.screen_size = tuple(int(v) for v in terrain.screen_size)
        H, W = self.screen_size

        self.U, self.U_dir = self._get_environment_parameters(environment)
        elevations = np.asarray(terrain.elevations, dtype=np.float64).reshape(H, W)
        self.slope_mag, self.slope_dir = compute_slopes(elevations, pixel_scale)  # fire.py:436-449
        w_0, delta, M_x, sigma = fuel_planes(np.asarray(terrain.fuels).reshape(H, W) if
                                             np.asarray(terrain.fuels).dtype == object else terrain.fuels)  # fmt: skip

        self._engine = FireEngine(
            H, W, 1, pixel_scale=pixel_scale, update_rate=update_rate, max_fire_duration=max_fire_duration,
            max_time=max_time, attenuate_line_ros=attenuate_line_ros, diagonal_spread=diagonal_spread,
            fuel_particle=(fuel_particle.h, fuel_particle.S_T, fuel_particle.S_e, fuel_particle.p_p),
            M_f=environment.M_f, keep_ros=keep_rate_of_spread, keep_ignition=keep_spread_graph, device=device,
        )  # fmt: skip
        self._engine.set_static(dict(w_0=w_0, delta=delta, M_x=M_x, sigma=sigma, U=self.U, U_dir=self.U_dir,
                                     slope_mag=self.slope_mag, slope_dir=self.slope_dir))  # fmt: skip
        self._engine.reset([self.init_pos])
        self._map8 = np.empty((H, W), dtype=np.int8)

    # fire.py:382-434 -- same accepted types, same errors
    def _get_environment_parameters(self, environment: Environment):
        def convert(param):
            if isinstance(param, float):
                return np.full(self.screen_size, param, dtype=np.float32)
            if isinstance(param, np.ndarray):
                if param.shape != self.screen_size:
                    raise ValueError(
                        f"The input parameter shape of {param.shape} should match the terrain shape "
                        f"of {self.screen_size}"
                    )
                return param
            if not all(isinstance(sub, collections.abc.Sequence) for sub in param):
                raise ValueError(
                    "The input parameter should be one of (float | Sequence[Sequence[float]] | "
                    f"np.ndarray), but got {type(param)}"
                )
            param = np.asarray(param)
            if param.shape != self.screen_size:
                raise ValueError(
                    f"The input parameter shape of {param.shape} should match the terrain shape "
                    f"of {self.screen_size}"
                )
            return param

        return convert(environment.U), convert(environment.U_dir)

    # -- the hot path ---------------------------------------------------------------------
    def update(self, fire_map: np.ndarray) -> Tuple[np.ndarray, GameStatus]:
        """
        One timestep (fire.py:616-719).  `fire_map` is the caller's (H, W) array of
        BurnStatus values: it is uploaded (the caller may have drawn control lines into it,
        mitigation.py:77), stepped on the device and updated in place, as the reference does.
        """
        if fire_map.shape != self.screen_size:
            raise AssertionError("The fire map does not match the shape of the terrain")  # fire.py:264-269
        np.copyto(self._map8, fire_map, casting="unsafe")
        status = self._engine.update(self._map8)
        np.copyto(fire_map, self._map8, casting="unsafe")
        return fire_map, GameStatus(int(status[0]))

    # -- attributes the reference exposes ---------------------------------------------------
    @property
    def elapsed_time(self) -> float:
        return float(self._engine.status()[1][0])

    @property
    def burn_amounts(self) -> np.ndarray:
        return self._engine.plane("burn")

    @property
    def rate_of_spread(self) -> np.ndarray:
        return self._engine.plane("ros")

    @property
    def durations(self):
        age = self._engine.plane("age")
        return [int(a) for a in age[age >= 0]]

    @property
    def sprites(self):
        """(x, y) of every cell that carries a Fire sprite, row-major."""
        ys, xs = np.nonzero(self._engine.plane("age") >= 0)
        return [(int(x), int(y)) for x, y in zip(xs, ys)]

    @property
    def fs_graph(self):
        """The fire-spread DiGraph the reference maintains per step (fire.py:380, :584), rebuilt
        from the device's ignition-step plane (needs `keep_spread_graph=True`)."""
        from .graph import to_networkx

        return to_networkx(self._engine.plane("ignition"), self.max_fire_duration)

    def spread_edges(self) -> np.ndarray:
        """Edges of that graph as an int32 [n, 4] array (x_src, y_src, x_dst, y_dst)."""
        from .graph import spread_edges

        return spread_edges(self._engine.plane("ignition"), self.max_fire_duration)

    @property
    def engine(self) -> FireEngine:
        return self._engine

    def close(self) -> None:
        self._engine.close()


__all__ = ["RothermelFireManager", "BurnStatus", "GameStatus"]


class ConstantSpreadFireManager:
    """
    Drop-in for `simfire.game.managers.fire.ConstantSpreadFireManager` (fire.py:722-787): same
    constructor, same `update(fire_map) -> fire_map`, run on the device through
    `sfb_constant_spread_update`.  The behaviour is the reference's as it executes, including its
    quirk: the Fire sprites that `update` appends have no duration entry, so the next call's
    `_prune_sprites` slices them off again (fire.py:148-155); the cells they marked stay BURNING
    and never spread.  The grid size is taken from the first `fire_map` (the reference's
    constructor has none either).  `sprites` lists the (x, y) of the sprites that still exist.
    """

    def __init__(self, init_pos: Tuple[int, int], fire_size: int, max_fire_duration: int, rate_of_spread: int,
                 *, device: int = 0) -> None:  # fmt: skip
        self.init_pos = tuple(int(v) for v in init_pos)
        self.fire_size = fire_size
        self.max_fire_duration = int(max_fire_duration)
        self.rate_of_spread = int(rate_of_spread)
        # FireManager.__init__ defaults (fire.py:49-103)
        self.attenuate_line_ros, self.headless, self.diagonal_spread = True, False, True
        self.device = device
        self._engine: Optional[FireEngine] = None
        self._calls = 0

    def _ensure_engine(self, shape) -> FireEngine:
        if self._engine is None:
            H, W = shape
            self._engine = FireEngine(H, W, 1, pixel_scale=1.0, update_rate=1.0, max_fire_duration=self.max_fire_duration,
                                      attenuate_line_ros=True, diagonal_spread=True, device=self.device)  # fmt: skip
            self._engine.reset([self.init_pos])  # the initial sprite (fire.py:101-103)
        elif (self._engine.H, self._engine.W) != tuple(shape):
            raise AssertionError(f"fire_map of shape {tuple(shape)}, the manager was started on {(self._engine.H, self._engine.W)}")
        return self._engine

    def update(self, fire_map: np.ndarray) -> np.ndarray:
        fm = np.asarray(fire_map)
        eng = self._ensure_engine(fm.shape)
        buf = np.ascontiguousarray(fm.astype(np.int8)).reshape(1, *fm.shape)
        eng.constant_spread_update(buf, self.rate_of_spread)
        self._calls += 1
        fire_map[...] = buf[0]  # the reference mutates the caller's array in place (fire.py:140, :779)
        return fire_map

    @property
    def sprites(self):
        # the initial sprite until it is pruned; the sprites appended by the spreading call live until the next call
        if self._engine is None:
            return [self.init_pos]
        age = self._engine.plane("age", 0)
        ys, xs = np.nonzero(age >= 0)
        return [(int(x), int(y)) for y, x in zip(ys, xs)]

    @property
    def durations(self):
        if self._engine is None:
            return [0]
        age = self._engine.plane("age", 0)
        return [int(a) for a in age[age >= 0]]

    def close(self) -> None:
        if self._engine is not None:
            self._engine.close()
            self._engine = None
