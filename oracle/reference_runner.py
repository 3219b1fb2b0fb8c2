"""
ORACLE SIDE (test / bench infrastructure, not product code): drive the UNMODIFIED reference.

`ReferenceFire` wraps one `RothermelFireManager` (simfire/game/managers/fire.py:287) of the staged
reference (`oracle/_ref`, made by `oracle/build_ref.py`; `/root/reference` in the dev container)
behind the same small interface as `oracle.dense_numpy.DenseFire` -- `step()`, `apply_points()`,
`status` (int8 fire_map), `elapsed_time`, `game_status` -- so that bench.py's CPU arm and the
same-run parity checks can use either.  The manager is constructed exactly the way
`FireSimulation._create_fire` does it (simulation.py:273-291): `terrain` is a stand-in object with
the three attributes the manager reads (`fuels`: object array of `Fuel`, `elevations`,
`screen_size`), `environment` a real `Environment`, `fire_map` an int64 (H, W) array
(simulation.py:561-566) that control lines are written into in place (mitigation.py:77).

Only `bench.py` (`--impl reference`, `cpu_baseline`) and `tests/` may import this module.
"""
from __future__ import annotations

import time
import types
from typing import Dict, Iterable, Optional, Tuple

import numpy as np

from . import ref_shim


def available() -> bool:
    return ref_shim.available()


def _fuel_array(mods, planes: Dict[str, np.ndarray], H: int, W: int) -> np.ndarray:
    """(H, W) object array of `Fuel` (simfire/world/parameters.py:31-50), one instance per distinct
    (w_0, delta, M_x, sigma) tuple -- `terrain.fuels` as the reference's layers build it."""
    rec = np.stack([np.broadcast_to(np.asarray(planes[k], dtype=np.float64), (H, W)) for k in ("w_0", "delta", "M_x", "sigma")], axis=-1)
    uniq, inv = np.unique(rec.reshape(-1, 4), axis=0, return_inverse=True)
    objs = np.empty(len(uniq), dtype=object)
    for i, (w_0, delta, M_x, sigma) in enumerate(uniq):
        objs[i] = mods.parameters.Fuel(w_0=float(w_0), delta=float(delta), M_x=float(M_x), sigma=float(sigma))
    return objs[inv.reshape(-1)].reshape(H, W)


class ReferenceFire:
    """One reference simulation (one env) on a window [y0:y0+h, x0:x0+w] of a workload's planes."""

    def __init__(self, planes: Dict[str, np.ndarray], elevations: Optional[np.ndarray], *, H: int, W: int,
                 start: Tuple[int, int], pixel_scale: float, update_rate: float, max_fire_duration: int,
                 max_time: Optional[float] = None, attenuate_line_ros: bool = True, diagonal_spread: bool = True,
                 M_f: float = 0.03, window: Optional[Tuple[int, int, int, int]] = None):  # fmt: skip
        fire, rothermel, enums, parameters, presets = ref_shim.import_reference()
        self._mods = types.SimpleNamespace(fire=fire, enums=enums, parameters=parameters)
        y0, x0, h, w = window if window is not None else (0, 0, H, W)
        self.window = (y0, x0, h, w)
        crop = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (H, W))[y0 : y0 + h, x0 : x0 + w])  # noqa: E731
        cp = {k: crop(v) for k, v in planes.items()}
        elev = crop(elevations if elevations is not None else 0.0)
        t0 = time.perf_counter()
        terrain = types.SimpleNamespace(fuels=_fuel_array(self._mods, cp, h, w), elevations=elev, screen_size=(h, w))
        env = parameters.Environment(float(M_f), cp["U"], cp["U_dir"])
        sx, sy = int(start[0]) - x0, int(start[1]) - y0
        self.mgr = fire.RothermelFireManager(
            (sx, sy), 2, int(max_fire_duration), pixel_scale, update_rate, parameters.FuelParticle(), terrain, env,
            max_time=max_time, attenuate_line_ros=bool(attenuate_line_ros), headless=True,
            diagonal_spread=bool(diagonal_spread),
        )  # fmt: skip
        self.init_seconds = time.perf_counter() - t0  # dominated by FireSpreadGraph (fire.py:380): one networkx node per pixel
        BS = enums.BurnStatus
        self.fire_map = np.full((h, w), BS.UNBURNED)  # int64, simulation.py:561-566
        self.fire_map[sy, sx] = BS.BURNING
        self.game_status = 1
        self.step_count = 0
        self._running = enums.GameStatus.RUNNING

    def apply_points(self, pts: Iterable[Tuple[int, int, int]]) -> None:
        y0, x0, h, w = self.window
        for x, y, kind in pts:
            if 0 <= y - y0 < h and 0 <= x - x0 < w:
                self.fire_map[y - y0, x - x0] = kind  # ControlLineManager.update, mitigation.py:77

    def step(self) -> int:
        """One `update()`; like `FireSimulation.run`, not called again once the status is QUIT."""
        if self.game_status != 1:
            return 0
        self.fire_map, st = self.mgr.update(self.fire_map)
        self.step_count += 1
        self.game_status = 1 if st == self._running else 0
        return self.game_status

    @property
    def status(self) -> np.ndarray:
        return self.fire_map.astype(np.int8)

    @property
    def elapsed_time(self) -> float:
        return float(self.mgr.elapsed_time)

    @property
    def burn(self) -> np.ndarray:
        return np.asarray(self.mgr.burn_amounts, dtype=np.float64)


def from_workload(wl, start, window=None) -> ReferenceFire:
    return ReferenceFire(wl.planes, wl.elevations, H=wl.H, W=wl.W, start=start, window=window, **wl.engine_kwargs())


def window_around(start, n_updates: int, H: int, W: int, margin: int = 3) -> Tuple[int, int, int, int]:
    """The (y0, x0, h, w) window a fire lit at `start` cannot leave in `n_updates` update() calls.

    A cell ignites only next to a burning cell, so after n calls every burning or burned cell is
    within Chebyshev distance n of the start and every candidate of call n within n + 1; `margin`
    more cells keep the window's own border (where np.gradient of a cropped elevation plane would
    differ) out of reach.  Inside the window a simulation of the crop is therefore identical, cell by
    cell, to the simulation of the full grid."""
    r = n_updates + 1 + margin
    x, y = int(start[0]), int(start[1])
    y0, y1 = max(0, y - r), min(H, y + r + 1)
    x0, x1 = max(0, x - r), min(W, x + r + 1)
    return (y0, x0, y1 - y0, x1 - x0)
