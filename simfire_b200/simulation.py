"""
`FireSimulation`: the SimHarness-facing surface of simfire/sim/simulation.py:184-553 --
`run()`, `fire_map`, `update_mitigation()`, `update_agent_positions()`, `load_mitigation()`,
`reset()`, `get_actions()`, `get_attribute_data()` ... -- on top of the device-resident
stepper, plus `BatchedFireSimulation`, the same surface vectorised over E independent envs
(one engine, one kernel launch per step for all of them).

Always headless: rendering, GIF recording and spread-graph drawing are display features
outside the hot-path scope.  `save_data`
(simulation.py:887-959) is kept: the per-update fire_map history is what downstream tooling reads.

Unlike the drop-in manager (`fire_manager.RothermelFireManager.update`, which round-trips the
caller's fire_map every call exactly like the reference), the simulation classes keep the map
on the device: mitigation goes down as a few (x, y, kind) points, `run(n)` issues n steps in
one call, and the int64 host `fire_map` is materialised lazily when it is read.
"""
from __future__ import annotations

import json
import warnings
from datetime import datetime
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from .config import Config, str_to_minutes
from .engine import FireEngine
from .enums import BurnStatus, ElevationConstants, FuelConstants, GameStatus, WindConstants
from .fire_manager import fuel_planes
from .parameters import Environment, FuelParticle


def _static_planes(config: Config):
    H, W = config.area.screen_size
    fdata = np.asarray(config.terrain.fuel_layer.data)
    fuels = fdata[..., 0] if fdata.dtype == object else fdata
    w_0, delta, M_x, sigma = fuel_planes(fuels)
    elev = np.asarray(config.terrain.topography_layer.data, dtype=np.float64).reshape(H, W)
    # slope planes are filled on the device by FireEngine.set_elevation
    return dict(w_0=w_0, delta=delta, M_x=M_x, sigma=sigma, U=config.wind.speed, U_dir=config.wind.direction,
                slope_mag=0.0, slope_dir=0.0), elev  # fmt: skip


def _engine_key(config: Config):
    """Everything `_engine_from_config` bakes into an engine: the reference builds a new
    RothermelFireManager from the CURRENT config on every reset() (simulation.py:273-290), and its config
    objects are mutable, so a reset must notice a change of any of these."""
    fp = FuelParticle()
    return (tuple(config.area.screen_size), float(config.area.pixel_scale), float(config.simulation.update_rate),
            int(config.fire.max_fire_duration), config.simulation.runtime, bool(config.mitigation.ros_attenuation),
            bool(config.fire.diagonal_spread), (fp.h, fp.S_T, fp.S_e, fp.p_p), float(config.environment.moisture))  # fmt: skip


def _engine_from_config(config: Config, E: int, device: int, shared_static: bool, **kw) -> FireEngine:
    fp = FuelParticle()
    H, W = config.area.screen_size
    return FireEngine(
        H, W, E, pixel_scale=config.area.pixel_scale, update_rate=config.simulation.update_rate,
        max_fire_duration=config.fire.max_fire_duration, max_time=config.simulation.runtime,
        attenuate_line_ros=config.mitigation.ros_attenuation, diagonal_spread=config.fire.diagonal_spread,
        fuel_particle=(fp.h, fp.S_T, fp.S_e, fp.p_p), M_f=config.environment.moisture,
        shared_static=shared_static, device=device, **kw,
    )  # fmt: skip


def _attribute_tensors(engine: FireEngine, elevations: np.ndarray, static_set: int):
    import torch

    rec = engine.static_device()[static_set]  # (H, W, 8) strided view, no copy
    return {"w_0": rec[..., 0], "sigma": rec[..., 3].to(torch.int32), "delta": rec[..., 1], "M_x": rec[..., 2],
            "elevation": torch.as_tensor(np.asarray(elevations), device=rec.device),
            "wind_speed": rec[..., 4], "wind_direction": rec[..., 5]}  # fmt: skip


class FireSimulation:
    def __init__(self, config: Config, *, device: int = 0) -> None:
        self.config = config
        self.device = device
        self._rendering = False
        self.agents: Dict[int, Tuple[int, int]] = {}
        self._engine: Optional[FireEngine] = None
        self.start_time = datetime.now().strftime("%Y-%m-%d_%H-%M-%S")  # simulation.py:56
        self.sf_home = Path(self.config.simulation.sf_home).expanduser()
        self.reset()

    # -- lifecycle (simulation.py:202-214) ----------------------------------------------------
    def reset(self) -> None:
        H, W = self.config.area.screen_size
        key = _engine_key(self.config)
        if self._engine is None or getattr(self, "_engine_key", None) != key:
            if self._engine is not None:
                self._engine.close()
            self._engine = _engine_from_config(self.config, 1, self.device, shared_static=True)
            self._engine_key = key
        self._planes, self._elevations = _static_planes(self.config)
        self._engine.set_static(self._planes)
        self._engine.set_elevation(self._elevations)  # slopes on the device (fire.py:436-449)
        self._engine.reset([self.config.fire.fire_initial_position])
        self.fuel_particle = FuelParticle()
        self.environment = Environment(self.config.environment.moisture, self.config.wind.speed,
                                       self.config.wind.direction)  # fmt: skip
        self.agents.clear()
        self.agent_positions = np.zeros((H, W), dtype=np.int64)
        self._fire_map: Optional[np.ndarray] = None
        self.elapsed_steps = 0
        self.elapsed_time = 0.0
        self.fire_status = GameStatus.RUNNING
        self.game_status = GameStatus.RUNNING
        self.active = True

    def close(self) -> None:
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    # -- fire_map: int64 (H, W) like the reference (simulation.py:561-564), fetched lazily -----
    @property
    def fire_map(self) -> np.ndarray:
        """int64 (H, W) copy of the device map, refreshed after every run() / mitigation call.  In the
        reference this is the live array and callers may write into it; here the state lives on the
        device, so the copy is READ-ONLY (an in-place write raises instead of being silently lost):
        assign a whole map (`sim.fire_map = m`, load_mitigation) or use update_mitigation()."""
        if self._fire_map is None:
            self._fire_map = self._engine.fire_map(0, 1)[0].astype(np.int64)
            self._fire_map.setflags(write=False)
        return self._fire_map

    @fire_map.setter
    def fire_map(self, value: np.ndarray) -> None:
        value = np.asarray(value)
        if value.shape != tuple(self.config.area.screen_size):
            raise ValueError(f"fire_map of shape {value.shape} does not match {self.config.area.screen_size}")
        self._engine.set_fire_map(value.astype(np.int8))
        self._fire_map = None

    def fire_map_device(self):
        """Zero-copy int8 (H, W) BurnStatus tensor on the GPU (observation for RL)."""
        return self._engine.fire_map_device()[0]

    # -- descriptive API ------------------------------------------------------------------------
    def get_actions(self) -> Dict[str, int]:
        return {"fireline": BurnStatus.FIRELINE, "scratchline": BurnStatus.SCRATCHLINE, "wetline": BurnStatus.WETLINE}

    @property
    def disaster_categories(self):
        return BurnStatus

    def get_disaster_categories(self) -> Dict[str, int]:
        return {i.name: i.value for i in BurnStatus}

    @staticmethod
    def supported_attributes() -> List[str]:
        return ["w_0", "sigma", "delta", "M_x", "elevation", "wind_speed", "wind_direction"]

    def get_attribute_bounds(self) -> Dict[str, object]:
        return {
            "w_0": {"min": FuelConstants.W_0_MIN, "max": FuelConstants.W_0_MAX},
            "sigma": {"min": FuelConstants.SIGMA_MIN, "max": FuelConstants.SIGMA_MAX},
            "delta": {"min": FuelConstants.DELTA_MIN, "max": FuelConstants.DELTA_MAX},
            "M_x": {"min": FuelConstants.M_X_MIN, "max": FuelConstants.M_X_MAX},
            "elevation": {"min": ElevationConstants.MIN_ELEVATION, "max": ElevationConstants.MAX_ELEVATION},
            "wind_speed": {"min": WindConstants.MIN_SPEED, "max": WindConstants.MAX_SPEED},
            "wind_direction": {"min": 0.0, "max": 360.0},
        }

    def get_attribute_data(self) -> Dict[str, np.ndarray]:
        """Same dtypes as simulation.py:395-403 (sigma is uint32 there)."""
        p = self._planes
        return {"w_0": p["w_0"].astype(np.float32), "sigma": p["sigma"].astype(np.uint32),
                "delta": p["delta"].astype(np.float32), "M_x": p["M_x"].astype(np.float32),
                # the layer's own array, dtype included (simulation.py:401: int64 for `flat`, whose function returns 0)
                "elevation": np.asarray(self.config.terrain.topography_layer.data).reshape(self._elevations.shape),
                "wind_speed": self.config.wind.speed,
                "wind_direction": self.config.wind.direction}  # fmt: skip

    def get_attribute_data_device(self):
        """`get_attribute_data` (simulation.py:376-403) as CUDA tensors for an on-device policy: zero-copy
        float32 (H, W) views of the stepper's own static records for the fuel and wind planes (sigma as int32:
        torch has no uint32 arithmetic; the reference's host dtype is uint32), elevation uploaded once."""
        return _attribute_tensors(self._engine, self._elevations, 0)

    # -- between-step mutations -------------------------------------------------------------------
    def load_mitigation(self, mitigation_map: np.ndarray) -> None:
        """simulation.py:425-447"""
        category_values = [status.value for status in BurnStatus]
        if np.isin(mitigation_map, category_values).all():
            message = ("You are overwriting the current fire map with the given mitigation map - the current "
                       "fire map data will be erased.")  # fmt: skip
            self.fire_map = mitigation_map
        else:
            message = f"Invalid values in {mitigation_map} - values need to be within {category_values}... Skipping"
        warnings.warn(message)

    def update_mitigation(self, points: Iterable[Tuple[int, int, int]]) -> None:
        """(column, row, mitigation) tuples (simulation.py:449-478); unknown kinds are skipped."""
        keep = [(0, int(c), int(r), int(m)) for c, r, m in points
                if m in (BurnStatus.FIRELINE, BurnStatus.SCRATCHLINE, BurnStatus.WETLINE)]  # fmt: skip
        if keep:
            self._engine.apply_points(keep)
            self._fire_map = None

    def update_agent_positions(self, points: Iterable[Tuple[int, int, int]]) -> None:
        """(column, row, agent_id) tuples (simulation.py:480-499); the fire never reads this."""
        for column, row, agent_id in points:
            self.agent_positions[self.agent_positions == agent_id] = 0
            self.agent_positions[row][column] = agent_id
            self.agents[agent_id] = (column, row)

    # -- run (simulation.py:501-553) -----------------------------------------------------------------
    def run(self, time: Union[str, int]) -> Tuple[np.ndarray, bool]:
        if isinstance(time, str):
            total_updates = round(str_to_minutes(time) / self.config.simulation.update_rate)
        else:
            total_updates = int(time)
        if self.fire_status == GameStatus.RUNNING and total_updates > 0 and self.config.simulation.save_data:
            # the history is written after every update (simulation.py:546-549): one step per launch
            for _ in range(total_updates):
                if self.fire_status != GameStatus.RUNNING:
                    break
                self._engine.step(1)
                st, el, n = self._engine.status()
                self.fire_status = GameStatus(int(st[0]))
                self.elapsed_time = float(el[0])
                self.elapsed_steps += 1  # counts update() calls, the one that returns QUIT included (:543)
                self._fire_map = None
                self._save_data()
        elif self.fire_status == GameStatus.RUNNING and total_updates > 0:
            # envs that return QUIT stop advancing on the device, like the `while` loop of the reference
            self._engine.step(total_updates)
            st, el, n = self._engine.status()
            self.fire_status = GameStatus(int(st[0]))
            self.elapsed_time = float(el[0])
            self.elapsed_steps = int(n[0])
            self._fire_map = None
        self.active = self.fire_status == GameStatus.RUNNING
        return self.fire_map, self.active

    # -- save_data (simulation.py:887-1106): metadata.json, the static layers once, and the fire_map
    # history under <sf_home>/data/<start_time>/ in the configured data type ---------------------------
    def _load_static_data(self, datapath: Path) -> Dict[str, object]:
        dtype = self.config.simulation.data_type
        data = self.get_attribute_data()
        ext = {"npy": "npy", "h5": "h5", "json": "json", "jsonl": "json"}.get(dtype)
        if ext is None:
            raise ValueError(f"Invalid data type '{dtype}' given. Valid types are 'npy', 'h5', 'json', and 'jsonl'.")
        data_locs = {k: f"{k}.{ext}" for k in data}
        shape = data[list(data.keys())[0]].shape
        for key, loc in data_locs.items():
            path = datapath / loc
            if path.is_file():
                continue
            if dtype == "npy":
                np.save(path, data[key])
            elif dtype == "h5":
                import h5py  # optional, like in the reference's environment

                with h5py.File(path, "w") as f:
                    f.create_dataset("data", data=data[key])
            else:
                with open(path, "w") as f:
                    json.dump({"data": data[key].tolist()}, f)
        return {"data": data_locs, "shape": shape}

    def _save_data(self) -> None:
        dtype = self.config.simulation.data_type
        ext = {"npy": "npy", "h5": "h5", "json": "jsonl", "jsonl": "jsonl"}.get(dtype)
        if ext is None:
            raise ValueError(f"Invalid data type '{dtype}' given. Valid types are 'npy', 'h5', 'json', and 'jsonl'.")
        datapath = self.sf_home / "data" / self.start_time
        datapath.mkdir(parents=True, exist_ok=True)
        fire_map_path = datapath / f"fire_map.{ext}"
        static = self._load_static_data(datapath)
        metadata = {"config": self.config.yaml_data, "seeds": self.get_seeds(), "layer_types": self.get_layer_types(),
                    "shape": static["shape"], "static_data": static, "fire_map": fire_map_path.name}  # fmt: skip
        with open(datapath / "metadata.json", "w") as f:
            json.dump(metadata, f, indent=2)
        current = np.expand_dims(self.fire_map, axis=0)
        if dtype in ("npy", "h5"):
            previous = None
            if fire_map_path.is_file():
                if dtype == "npy":
                    previous = np.load(fire_map_path)
                else:
                    import h5py

                    previous = np.array(h5py.File(fire_map_path)["data"])
                if previous.ndim == 2:
                    previous = np.expand_dims(previous, axis=0)
            history = current if previous is None else np.append(previous, current, axis=0)
            if dtype == "npy":
                np.save(fire_map_path, history.astype(np.int8))
            else:
                import h5py

                with h5py.File(fire_map_path, "w") as f:
                    f.create_dataset("data", data=history)
        else:  # JSON Lines: one {elapsed_steps: fire_map} object per update
            with open(fire_map_path, "a") as f:
                f.write(json.dumps({self.elapsed_steps: self.fire_map.tolist()}) + "\n")

    # -- config mutation helpers the harness uses (simulation.py:574-829); like the reference they
    # only rewrite `self.config`: the change takes effect at the next reset() ----------------------------
    def set_fire_initial_position(self, pos: Tuple[int, int]) -> None:
        self.config.reset_fire(pos=pos)

    def get_seeds(self) -> Dict[str, Optional[int]]:
        """Seeds that exist for the configured layers; None-valued ones are left out (simulation.py:574-597)."""
        seeds = {"elevation": self._get_topography_seed(), "fuel": self._get_fuel_seed(),
                 "wind_speed": self._function_seed(self.config.wind.speed_function),
                 "wind_direction": self._function_seed(self.config.wind.direction_function),
                 "fire_initial_position": self.config.fire.seed}  # fmt: skip
        return {k: v for k, v in seeds.items() if v is not None}

    @staticmethod
    def _function_seed(fn) -> Optional[int]:
        return fn.kwargs["seed"] if fn is not None and fn.name == "perlin" else None

    def _get_topography_seed(self) -> Optional[int]:
        t = self.config.terrain
        if t.topography_type != "functional":
            return None  # array-backed layers: nothing to re-seed
        if t.topography_function is None:
            raise RuntimeError("The topography type is set as functional, but "
                               "self.config.terrain.topography_function is not set")  # fmt: skip
        if t.topography_function.name == "perlin":
            return t.topography_function.kwargs["seed"]
        if t.topography_function.name in ("flat", "gaussian"):  # the reference rejects 'gaussian' here (:617-622)
            return None
        raise RuntimeError(f"The topography function name {t.topography_function.name} is not valid")

    def _get_fuel_seed(self) -> Optional[int]:
        t = self.config.terrain
        if t.fuel_type != "functional":
            return None
        if t.fuel_function is None:
            raise RuntimeError("The fuel type is set as functional, but self.config.terrain.fuel_function is not set")
        if t.fuel_function.name == "chaparral":
            return t.fuel_function.kwargs["seed"]
        raise RuntimeError(f"The fuel function name {t.fuel_function.name} is not valid")

    def set_seeds(self, seeds: Dict[str, int]) -> bool:
        """simulation.py:713-759: every recognised key is applied; any unknown key makes the
        call report failure (after the known ones were applied, as in the reference)."""
        success = False
        if "elevation" in seeds:
            self.config.reset_terrain(topography_seed=seeds["elevation"])
            success = True
        if "fuel" in seeds:
            self.config.reset_terrain(fuel_seed=seeds["fuel"])
            success = True
        if "wind_speed" in seeds or "wind_direction" in seeds:
            self.config.reset_wind(speed_seed=seeds.get("wind_speed"), direction_seed=seeds.get("wind_direction"))
            success = True
        if "fire_initial_position" in seeds:
            self.config.reset_fire(seeds["fire_initial_position"])
        valid_keys = list(self.get_seeds().keys())
        for key in seeds:
            if key not in valid_keys:
                warnings.warn("No valid keys in the seeds dictionary were given to the set_seeds method. No seeds "
                              f"will be changed. Valid keys are: {valid_keys}")  # fmt: skip
                success = False
        return success

    def get_layer_types(self) -> Dict[str, str]:
        return {"elevation": self.config.terrain.topography_type, "fuel": self.config.terrain.fuel_type}

    def set_layer_types(self, types: Dict[str, str]) -> bool:
        """simulation.py:784-829.  Only 'functional' layers are generated here; asking for
        'operational' raises ConfigError from Config.reset_terrain (LANDFIRE ingest is out of scope)."""
        valid_keys = list(self.get_layer_types().keys())
        bad = [k for k in types if k not in valid_keys]
        if bad or not types:
            if bad:
                warnings.warn("No valid keys in the types dictionary were given to the set_data_types method. No "
                              f"data types will be changed. Valid keys are: {valid_keys}")  # fmt: skip
            return False
        self.config.reset_terrain(topography_type=types.get("elevation"), fuel_type=types.get("fuel"))
        return True

    def rendering(self) -> bool:
        return False


class BatchedFireSimulation:
    """
    E independent `FireSimulation`s of one terrain on one GPU.  Methods mirror
    `FireSimulation` with a leading env axis: `run(n)` -> (fire_maps[E, H, W] int8, active[E]),
    `update_mitigation(points)` takes (env, column, row, mitigation) rows, `reset(envs=...)`
    restarts a subset (the RL "done" handling).  Across GPUs, run one instance per process /
    device on a disjoint slice of envs: envs never interact, so there is no collective.
    """

    def __init__(self, config: Config, num_envs: int, *, device: int = 0,
                 initial_positions: Optional[Sequence[Tuple[int, int]]] = None) -> None:  # fmt: skip
        self.config = config
        self.num_envs = int(num_envs)
        self._engine = _engine_from_config(config, self.num_envs, device, shared_static=True, track_changes=True)
        self._planes, self._elevations = _static_planes(config)
        self._mirror: Optional[np.ndarray] = None
        self._engine.set_static(self._planes)
        self._engine.set_elevation(self._elevations)  # slopes on the device (fire.py:436-449)
        pos = initial_positions if initial_positions is not None else [config.fire.fire_initial_position] * self.num_envs
        self._engine.reset(pos)
        H, W = config.area.screen_size
        self.agent_positions = np.zeros((self.num_envs, H, W), dtype=np.int64)
        self._agents_dev = None

    @property
    def engine(self) -> FireEngine:
        return self._engine

    def reset(self, positions=None, envs: Optional[Sequence[int]] = None) -> None:
        envs = list(range(self.num_envs)) if envs is None else list(envs)
        if positions is None:
            positions = [self.config.fire.fire_initial_position] * len(envs)
        self._engine.reset(positions, envs=envs)
        self.agent_positions[envs] = 0
        if self._agents_dev is not None:
            self._agents_dev[envs] = 0

    def update_mitigation(self, points) -> None:
        pts = np.asarray(points, dtype=np.int32).reshape(-1, 4)
        ok = (pts[:, 3] >= BurnStatus.FIRELINE) & (pts[:, 3] <= BurnStatus.WETLINE)
        self._engine.apply_points(pts[ok])

    def update_agent_positions(self, points) -> None:
        for env, column, row, agent_id in points:
            a = self.agent_positions[env]
            old = np.argwhere(a == agent_id)
            a[a == agent_id] = 0
            a[row, column] = agent_id
            if self._agents_dev is not None:  # a few scalar writes per agent, no plane crosses PCIe
                for oy, ox in old:
                    self._agents_dev[env, int(oy), int(ox)] = 0
                self._agents_dev[env, row, column] = agent_id

    def run(self, time: Union[str, int], out: Optional[np.ndarray] = None):
        if isinstance(time, str):
            total_updates = round(str_to_minutes(time) / self.config.simulation.update_rate)
        else:
            total_updates = int(time)
        self._engine.step(total_updates, sync=False)
        if out is None:
            if self._mirror is None:
                H, W = self.config.area.screen_size
                self._mirror = np.empty((self.num_envs, H, W), dtype=np.int8)
            out = self._mirror
        # the same array is returned (updated in place) on every call, like the reference's
        # fire_map; only the cells that changed since the previous call cross PCIe
        self._engine.sync_fire_maps(out)
        st, _, _ = self._engine.status()
        return out, st == GameStatus.RUNNING

    def fire_maps_device(self):
        return self._engine.fire_map_device()

    def get_attribute_data_device(self):
        """The shared terrain's attribute planes as CUDA tensors (see FireSimulation.get_attribute_data_device)."""
        return _attribute_tensors(self._engine, self._elevations, 0)

    @property
    def agent_positions_device(self):
        """`agent_positions` (simulation.py:480-499) as an int16 CUDA tensor [E, H, W], kept up to date by
        update_agent_positions from the first access on (RL observation without a host round trip)."""
        import torch

        if self._agents_dev is None:
            self._agents_dev = torch.as_tensor(self.agent_positions.astype(np.int16), device=f"cuda:{self._engine.device}")
        return self._agents_dev

    @property
    def elapsed_time(self) -> np.ndarray:
        return self._engine.status()[1]

    @property
    def elapsed_steps(self) -> np.ndarray:
        return self._engine.status()[2]

    def close(self) -> None:
        self._engine.close()
