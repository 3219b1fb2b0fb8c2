"""
Synthetic inputs of the shapes BASELINE.json names and the fuel
tables needed to build them.  Host-side NumPy, run once per benchmark / test: this is the
"data layer" side of the boundary (simfire/utils/layers.py in the reference), not the hot
path.  Every generator returns a `Workload` whose `planes` are the eight (H, W) static
inputs of the stepper and whose scalar fields are RothermelFireManager's constructor
arguments (fire.py:293-307).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

STATIC_PLANES = ("w_0", "delta", "M_x", "sigma", "U", "U_dir", "slope_mag", "slope_dir")

# Anderson-13 fire behaviour fuel models + non-burnable classes as (w_0, delta, M_x, sigma):
# the values SimFire ships in simfire/world/presets.py:17-55, keyed by the LANDFIRE fuel
# model id as in `FuelModelToFuel` (simfire/enums.py:176-198).
FUEL_MODELS: Dict[int, Tuple[float, float, float, float]] = {
    1: (0.0340, 1.000, 0.1200, 3500.0),   # ShortGrass
    2: (0.0918, 1.000, 0.1500, 2784.0),   # GrassTimberShrubOverstory
    3: (0.1377, 2.500, 0.2500, 1500.0),   # TallGrass
    4: (0.2296, 6.000, 0.2000, 1739.0),   # Chaparral
    5: (0.0459, 2.000, 0.2000, 1683.0),   # Brush
    6: (0.0688, 2.500, 0.2500, 1564.0),   # DormantBrushHardwoodSlash
    7: (0.0459, 2.500, 0.4000, 1552.0),   # SouthernRough
    8: (0.0688, 0.200, 0.3000, 1889.0),   # ClosedShortNeedleTimberLitter
    9: (0.1331, 0.200, 0.2500, 2484.0),   # HardwoodLongNeedlePineTimber
    10: (0.1377, 1.000, 0.2500, 1764.0),  # TimberLitterUnderstory
    11: (0.0688, 1.000, 0.1500, 1182.0),  # LightLoggingSlash
    12: (0.1836, 2.300, 0.2000, 1145.0),  # MediumLoggingSlash
    13: (0.3214, 3.000, 0.2500, 1159.0),  # HeavyLoggingSlash
    91: (0.0, 1.0, 1.0, 1.0),             # NBUrban
    92: (0.0, 1.0, 1.0, 1.0),             # NBSnowIce
    93: (0.0, 1.0, 1.0, 1.0),             # NBAgriculture
    98: (0.0, 1.0, 1.0, 1.0),             # NBWater
    99: (0.0, 1.0, 1.0, 1.0),             # NBBarren
}  # fmt: skip
BURNABLE_IDS = tuple(range(1, 14))
NONBURNABLE_IDS = (91, 92, 93, 98, 99)

MPH_TO_FTPM = 88.0  # simfire/utils/units.py:34


def fuel_planes_from_ids(ids: np.ndarray) -> Dict[str, np.ndarray]:
    """(H, W) fuel-model ids -> w_0 / delta / M_x / sigma planes (float64)."""
    lut = np.zeros((100, 4))
    for k, v in FUEL_MODELS.items():
        lut[k] = v
    rec = lut[ids]
    return {"w_0": rec[..., 0].copy(), "delta": rec[..., 1].copy(), "M_x": rec[..., 2].copy(), "sigma": rec[..., 3].copy()}


def compute_slopes(elevations: np.ndarray, pixel_scale: float):
    """`RothermelFireManager._compute_slopes` (fire.py:436-449): float64 gradient planes."""
    grad_y, grad_x = np.gradient(np.asarray(elevations, dtype=np.float64), pixel_scale)
    return np.sqrt(grad_x**2 + grad_y**2), np.arctan2(grad_y, grad_x + 0.000001)


@dataclass
class Workload:
    name: str
    H: int
    W: int
    planes: Dict[str, np.ndarray]
    pixel_scale: float
    update_rate: float = 1.0
    max_fire_duration: int = 4
    max_time: Optional[float] = None
    attenuate_line_ros: bool = True
    diagonal_spread: bool = True
    M_f: float = 0.03
    init_pos: Tuple[int, int] = (0, 0)  # (x, y)
    description: str = ""
    elevations: Optional[np.ndarray] = field(default=None, repr=False)

    def engine_kwargs(self) -> dict:
        return dict(pixel_scale=self.pixel_scale, update_rate=self.update_rate,
                    max_fire_duration=self.max_fire_duration, max_time=self.max_time,
                    attenuate_line_ros=self.attenuate_line_ros, diagonal_spread=self.diagonal_spread,
                    M_f=self.M_f)  # fmt: skip

    def burnable_starts(self, n: int, seed: int, margin: int = 8) -> np.ndarray:
        """n random (x, y) ignition cells on burnable fuel, away from the border."""
        rng = np.random.default_rng(seed)
        w0 = self.planes["w_0"]
        out = np.empty((n, 2), dtype=np.int32)
        k = 0
        while k < n:
            x = rng.integers(margin, self.W - margin, n)
            y = rng.integers(margin, self.H - margin, n)
            ok = np.broadcast_to(w0, (self.H, self.W))[y, x] > 0
            m = min(int(ok.sum()), n - k)
            out[k : k + m, 0] = x[ok][:m]
            out[k : k + m, 1] = y[ok][:m]
            k += m
        return out


def cfg1_functional_flat(size: int = 128, start: Optional[Tuple[int, int]] = None) -> Workload:
    """
    BASELINE config 1: `configs/functional_config.yml` with screen_size [128, 128] and flat
    topography: one fuel everywhere (chaparral(seed=1113), values quoted in SURVEY.md 8c-i),
    simple wind 7 mph @ 90 deg, pixel_scale 50, update_rate 1, max_fire_duration 4,
    8-neighbour, attenuation on, moisture 0.03, runtime 24 h.
    """
    H = W = size
    planes = dict(w_0=0.9810356625846572, delta=5.890006842991012, M_x=0.9833113830744984,
                  sigma=3433.643783383716, U=7.0 * MPH_TO_FTPM, U_dir=90.0, slope_mag=0.0, slope_dir=0.0)  # fmt: skip
    planes = {k: np.full((H, W), v, dtype=np.float64) for k, v in planes.items()}
    planes["slope_mag"], planes["slope_dir"] = compute_slopes(np.zeros((H, W)), 50.0)
    return Workload("cfg1_functional_flat_%d" % size, H, W, planes, pixel_scale=50.0, update_rate=1.0,
                    max_fire_duration=4, max_time=1440.0, attenuate_line_ros=True, diagonal_spread=True,
                    M_f=0.03, init_pos=start or (W // 2, H // 2),
                    description="functional_config.yml, flat, chaparral(seed=1113), wind 616 ft/min @ 90")  # fmt: skip


def _smooth_field(rng, H: int, W: int, lo: float, hi: float, k: int = 4) -> np.ndarray:
    """Sum of a few low-frequency cosine products, rescaled to [lo, hi]."""
    y = np.arange(H, dtype=np.float64)[:, None] / H
    x = np.arange(W, dtype=np.float64)[None, :] / W
    f = np.zeros((H, W))
    for _ in range(k):
        fy, fx = rng.uniform(0.3, 2.5, 2)
        py, px = rng.uniform(0, 2 * np.pi, 2)
        f += rng.uniform(0.3, 1.0) * np.cos(2 * np.pi * fy * y + py) * np.cos(2 * np.pi * fx * x + px)
    f -= f.min()
    f /= max(float(f.max()), 1e-12)
    return lo + (hi - lo) * f


def synthetic_operational(H: int, W: int, seed: int = 0, *, flat: bool = False, patch: int = 32,
                          name: Optional[str] = None) -> Workload:  # fmt: skip
    """
    BASELINE configs 2-5 ("synthetic operational terrain"): fuel-model id per
    patch x patch block (85 % Anderson-13, 15 % non-burnable), elevation = four Gaussian
    bumps in 0..3000 ft (flat=True: zero), wind speed 7..47 mph and direction 0..360 deg as
    smooth fields, pixel_scale = int(30 m in ft) = 98, moisture 0.001, max_fire_duration 5,
    no attenuation, 8-neighbour, update_rate 1; ignition at the centre, forced burnable.
    """
    rng = np.random.default_rng(seed)
    ph, pw = -(-H // patch), -(-W // patch)
    ids = np.where(rng.random((ph, pw)) < 0.85, rng.choice(BURNABLE_IDS, (ph, pw)), rng.choice(NONBURNABLE_IDS, (ph, pw)))
    ids = np.repeat(np.repeat(ids, patch, axis=0), patch, axis=1)[:H, :W]
    x0, y0 = W // 2, H // 2
    if FUEL_MODELS[int(ids[y0, x0])][0] <= 0:
        ids[y0 - y0 % patch : y0 - y0 % patch + patch, x0 - x0 % patch : x0 - x0 % patch + patch] = 4
    planes = fuel_planes_from_ids(ids)
    if flat:
        elev = np.zeros((H, W))
    else:
        yy = np.arange(H, dtype=np.float64)[:, None]
        xx = np.arange(W, dtype=np.float64)[None, :]
        elev = np.zeros((H, W))
        for _ in range(4):
            cy, cx = rng.uniform(0, H), rng.uniform(0, W)
            sy, sx = rng.uniform(0.1, 0.35) * H, rng.uniform(0.1, 0.35) * W
            elev += rng.uniform(0.4, 1.0) * np.exp(-(((yy - cy) / sy) ** 2 + ((xx - cx) / sx) ** 2))
        elev = 3000.0 * (elev - elev.min()) / max(float(elev.max() - elev.min()), 1e-12)
    ps = float(int(30 / 0.3048))  # simfire/utils/config.py:482-484
    planes["U"] = _smooth_field(rng, H, W, 7.0, 47.0) * MPH_TO_FTPM
    planes["U_dir"] = _smooth_field(rng, H, W, 0.0, 360.0)
    planes["slope_mag"], planes["slope_dir"] = compute_slopes(elev, ps)
    return Workload(name or f"synthetic_operational_{H}x{W}_seed{seed}{'_flat' if flat else ''}", H, W, planes,
                    pixel_scale=ps, update_rate=1.0, max_fire_duration=5, max_time=None,
                    attenuate_line_ros=False, diagonal_spread=True, M_f=0.001, init_pos=(x0, y0),
                    description=f"fuel models per {patch}x{patch} patch (85% Anderson-13, 15% non-burnable), "
                                f"{'flat' if flat else '4 Gaussian hills 0-3000 ft'}, wind 7-47 mph smooth field",
                    elevations=elev)  # fmt: skip
