"""
ORACLE (test infrastructure, not product code) -- dense per-cell restatement of one
``RothermelFireManager.update`` timestep (``/root/reference/simfire/game/managers/fire.py:616-719``).

The reference walks a Python list of burning "sprites"; this restates the same
semantics as whole-grid NumPy operations on four planes per simulation:

    status  int8   BurnStatus value 0..5 (``simfire/enums.py:52-69``)
    live    bool   a Fire sprite exists on the cell (``fire.py:101-103``, ``:571-579``)
    ign     int32  update() call (1-based) that ignited the sprite; 0 = initial fire
    burn    f64    ``burn_amounts`` (``fire.py:370``, accumulated at ``:710``)

It is the executable specification of the CUDA kernel.  Parity with the reference
itself (fire_map, burn_amounts and rate_of_spread, every step) is pinned by
``tests/test_oracle_numpy.py`` against golden trajectories produced from the unmodified
reference by ``tests/golden/gen_golden.py``.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu-baseline leg may import this module.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .rothermel_numpy import NEIGHBOUR_DIRS_4, NEIGHBOUR_OFFSETS, rate_of_spread

UNBURNED, BURNING, BURNED, FIRELINE, SCRATCHLINE, WETLINE = range(6)
RUNNING, QUIT = 1, 0
# RoSAttenuation, simfire/enums.py:83-85, indexed by BurnStatus
ATTENUATION = np.array([0.0, 0.0, 0.0, 980.0, 490.0, 245.0])

STATIC_PLANES = ("w_0", "delta", "M_x", "sigma", "U", "U_dir", "slope_mag", "slope_dir")


def compute_slopes(elevations: np.ndarray, pixel_scale: float):
    """``RothermelFireManager._compute_slopes`` (fire.py:436-449), float64."""
    grad_y, grad_x = np.gradient(np.asarray(elevations, dtype=np.float64), pixel_scale)
    return np.sqrt(grad_x**2 + grad_y**2), np.arctan2(grad_y, grad_x + 0.000001)


@dataclass
class DenseParams:
    pixel_scale: float = 50.0
    update_rate: float = 1.0
    max_fire_duration: int = 4
    max_time: Optional[float] = None
    attenuate_line_ros: bool = True
    diagonal_spread: bool = True
    h: float = 8000.0
    S_T: float = 0.0555
    S_e: float = 0.01
    p_p: float = 32.0
    M_f: float = 0.03


@dataclass
class DenseFire:
    """One simulation instance (one env)."""

    planes: dict  # 8 static (H, W) arrays, any float dtype (cast to f32 at use, fire.py:537)
    params: DenseParams
    init_pos: tuple  # (x, y) as in the reference
    status: np.ndarray = field(init=False)
    live: np.ndarray = field(init=False)
    ign: np.ndarray = field(init=False)
    burn: np.ndarray = field(init=False)
    ros: np.ndarray = field(init=False)
    step_count: int = 0  # number of update() calls made so far
    elapsed_time: float = 0.0
    game_status: int = RUNNING

    def __post_init__(self):
        shape = np.asarray(self.planes["w_0"]).shape
        self.H, self.W = shape
        self.status = np.zeros(shape, dtype=np.int8)
        self.live = np.zeros(shape, dtype=bool)
        self.ign = np.full(shape, -1, dtype=np.int32)
        self.burn = np.zeros(shape, dtype=np.float64)
        self.ros = np.zeros(shape, dtype=np.float64)
        x0, y0 = self.init_pos
        self.status[y0, x0] = BURNING  # simulation.py:561-566
        self.live[y0, x0] = True  # fire.py:101-103
        self.ign[y0, x0] = 0

    # -- between-step mutations -------------------------------------------------------
    def apply_points(self, points):
        """``ControlLineManager.update`` (mitigation.py:60-80): unconditional overwrite."""
        for x, y, kind in points:
            self.status[y, x] = kind

    def set_fire_map(self, fire_map: np.ndarray):
        self.status[...] = np.asarray(fire_map, dtype=np.int8)

    # -- the timestep -----------------------------------------------------------------
    def step(self) -> int:
        p = self.params
        if self.game_status != RUNNING:  # simulation.py:533 never calls update() again
            return self.game_status
        self.step_count += 1
        t = self.step_count

        # 1. prune (fire.py:116-161): duration seen at step t is t-1-ign
        expired = self.live & ((t - 1 - self.ign) >= p.max_fire_duration)
        self.status[expired] = BURNED
        self.live[expired] = False
        # 2. no sprites -> QUIT (fire.py:637)
        if not self.live.any():
            self.game_status = QUIT
            return QUIT
        # 3. end of simulated time -> QUIT (fire.py:641-643)
        if p.max_time is not None and (
            p.update_rate > p.max_time or self.elapsed_time > p.max_time
        ):
            self.game_status = QUIT
            return QUIT

        # 4. candidates and the sprite whose pair is written last (fire.py:163-234, :704-705)
        H, W = self.H, self.W
        ignitable = (self.status == UNBURNED) | (self.status >= FIRELINE)
        dirs = range(8) if p.diagonal_spread else NEIGHBOUR_DIRS_4
        yy, xx = np.mgrid[0:H, 0:W]
        best_key = np.full((H, W), -1, dtype=np.int64)
        best_dir = np.zeros((H, W), dtype=np.int8)
        for k in dirs:
            dx, dy = NEIGHBOUR_OFFSETS[k]
            ys, xs = yy - dy, xx - dx  # source cell of a pair travelling in direction k
            ok = (ys >= 0) & (ys < H) & (xs >= 0) & (xs < W)
            ysc, xsc = np.clip(ys, 0, H - 1), np.clip(xs, 0, W - 1)
            ok &= self.live[ysc, xsc]
            # sprite-list order = (ignition step, y, x); the last pair written wins
            key = (self.ign[ysc, xsc].astype(np.int64) * H + ysc) * W + xsc
            better = ok & (key > best_key)
            best_key = np.where(better, key, best_key)
            best_dir = np.where(better, np.int8(k), best_dir)
        cand = ignitable & (best_key >= 0)
        # 5. sprites but nowhere to go: nothing else happens (fire.py:651-652)
        if not cand.any():
            return RUNNING

        # 6. rate of spread of the winning pair, destination cell's parameters
        cy, cx = np.nonzero(cand)
        args = [np.asarray(self.planes[name])[cy, cx] for name in STATIC_PLANES]
        R = rate_of_spread(
            best_dir[cy, cx], *args, h=p.h, S_T=p.S_T, S_e=p.S_e, p_p=p.p_p, M_f=p.M_f
        )
        R = R * p.update_rate  # fire.py:696
        ros = np.zeros((H, W), dtype=np.float64)
        ros[cy, cx] = R
        # 7. control lines (fire.py:236-284): on EVERY line cell of the map
        line = self.status >= FIRELINE
        if p.attenuate_line_ros:
            ros = ros - np.where(line, ATTENUATION[self.status], 0.0)
        else:
            ros[line] = 0.0
        self.ros = ros
        # 8. accumulate (fire.py:710)
        self.burn = self.burn + ros
        # 9. ignite (fire.py:566-587): strict '>' on the accumulated burn
        new = cand & (self.burn > p.pixel_scale)
        self.status[new] = BURNING
        self.live[new] = True
        self.ign[new] = t
        # 10. fire.py:717
        self.elapsed_time += p.update_rate
        return RUNNING

    @property
    def fire_map(self) -> np.ndarray:
        return self.status.astype(np.int64)
