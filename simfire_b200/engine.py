"""
FireEngine: thin object wrapper over the C ABI (include/simfire_b200.h).  One engine = one
CUDA device = E independent fire simulations of one H x W grid, all device-resident.

This is the layer the drop-in manager (`simfire_b200.fire_manager.RothermelFireManager`)
and the batched RL surface (`simfire_b200.simulation`) are written on.  It holds no
arithmetic of its own: every method is one call into libsimfire_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Optional, Sequence, Union

import numpy as np

from . import _lib
from ._lib import STATIC_PLANES

ArrayOrFloat = Union[np.ndarray, float]


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class FireEngine:
    """
    Device-resident state of E `RothermelFireManager`s (simfire/game/managers/fire.py:287).

    Constructor arguments mirror `RothermelFireManager.__init__` (fire.py:293-307):
    pixel_scale, update_rate, max_fire_duration, max_time, attenuate_line_ros,
    diagonal_spread; `fuel_particle` is (h, S_T, S_e, p_p) of `FuelParticle`
    (simfire/world/parameters.py:8-27) and `M_f` is `Environment.M_f` (:53).
    """

    def __init__(
        self,
        H: int,
        W: int,
        E: int = 1,
        *,
        pixel_scale: float,
        update_rate: float,
        max_fire_duration: int,
        max_time: Optional[float] = None,
        attenuate_line_ros: bool = True,
        diagonal_spread: bool = True,
        fuel_particle: Sequence[float] = (8000.0, 0.0555, 0.01, 32.0),
        M_f: float = 0.03,
        shared_static: bool = False,
        keep_ros: bool = False,
        device: int = 0,
        rows_per_chunk: int = 0,
        queue_capacity: int = 0,
        wide_cells: bool = False,
        sweep_ldg: bool = False,
        track_changes: bool = False,
        keep_ignition: bool = False,
        env_groups: int = 0,
        unit_skip: Optional[bool] = None,
        unit_chunks: bool = False,
        step_graph: bool = False,
        front_lists: bool = False,
        front_bits: bool = False,
        slab_y0: int = 0,
        slab_total_H: int = 0,
    ) -> None:
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.H, self.W, self.E = int(H), int(W), int(E)
        self.device = int(device)
        self.shared_static = bool(shared_static)
        self.keep_ros = bool(keep_ros)
        flags = 0
        flags |= _lib.DIAGONAL_SPREAD if diagonal_spread else 0
        flags |= _lib.ATTENUATE_LINE_ROS if attenuate_line_ros else 0
        flags |= _lib.SHARED_STATIC if shared_static else 0
        flags |= _lib.KEEP_ROS if keep_ros else 0
        flags |= _lib.HAS_MAX_TIME if max_time is not None else 0
        flags |= _lib.WIDE_CELLS if wide_cells else 0
        flags |= _lib.SWEEP_LDG if sweep_ldg else 0
        flags |= _lib.TRACK_CHANGES if track_changes else 0
        flags |= _lib.KEEP_IGNITION if keep_ignition else 0
        # None: the library decides (on from 1024 sweep units up); results do not depend on it
        if unit_skip is not None:
            flags |= _lib.UNIT_SKIP_ON if unit_skip else _lib.UNIT_SKIP_OFF
        flags |= _lib.UNIT_CHUNKS if unit_chunks else 0  # chunk-of-rows units + sweep instead of row units
        flags |= _lib.STEP_GRAPH if step_graph else 0    # multi-group handles: pairs of steps as one CUDA graph
        flags |= _lib.FRONT_LISTS if front_lists else 0  # the list-driven step (k_front)
        flags |= _lib.FRONT_BITS if front_bits else 0    # the bitboard front end (k_tiles); needs max_fire_duration <= 7
        h, S_T, S_e, p_p = (float(v) for v in fuel_particle)
        prm = _lib.SfbParams(
            abi_version=_lib.ABI_VERSION, device=self.device, H=self.H, W=self.W, E=self.E,
            max_fire_duration=int(max_fire_duration), flags=flags, rows_per_chunk=int(rows_per_chunk),
            pixel_scale=float(pixel_scale), update_rate=float(update_rate),
            max_time=float(max_time) if max_time is not None else 0.0,
            h=h, S_T=S_T, S_e=S_e, p_p=p_p, M_f=float(M_f), env_groups=int(env_groups),
            queue_capacity=int(queue_capacity), slab_y0=int(slab_y0), slab_total_H=int(slab_total_H),
        )  # fmt: skip
        _lib.check(self._lib.sfb_create(C.byref(prm), C.byref(self._h)))

    # -- lifetime -------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.sfb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- static planes --------------------------------------------------------------------
    def set_static(self, planes: Mapping[str, ArrayOrFloat], env: int = -1) -> None:
        """
        Upload w_0, delta, M_x, sigma, U, U_dir, slope_mag, slope_dir ((H, W) arrays or
        scalars) for `env` (-1: all envs / the shared set).  Values are cast to float32
        here, which is where the reference casts them too (fire.py:537, :546).
        """
        buf = np.empty((len(STATIC_PLANES), self.H, self.W), dtype=np.float32)
        for i, name in enumerate(STATIC_PLANES):
            a = np.asarray(planes[name], dtype=np.float64)
            if a.ndim != 0 and a.shape != (self.H, self.W):
                raise ValueError(
                    f"The input parameter shape of {a.shape} should match the terrain shape of {(self.H, self.W)}"
                )
            buf[i] = a.astype(np.float32)
        _lib.check(self._lib.sfb_set_static_all(self._h, int(env), _ptr(buf)))

    def set_elevation(self, elevations: np.ndarray, env: int = -1) -> None:
        """`_compute_slopes` (fire.py:436-449) on the device: fills the slope_mag / slope_dir planes
        of `env` from (H, W) elevations in ft.  Call after `set_static` (which overwrites them)."""
        e = np.ascontiguousarray(np.asarray(elevations, dtype=np.float64).reshape(self.H, self.W))
        _lib.check(self._lib.sfb_set_elevation(self._h, int(env), _ptr(e)))

    def set_static_plane(self, name: str, values: ArrayOrFloat, env: int = -1) -> None:
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(values, dtype=np.float64), (self.H, self.W)).astype(np.float32))
        _lib.check(self._lib.sfb_set_static(self._h, int(env), STATIC_PLANES.index(name), _ptr(a)))

    # -- between-step mutations -----------------------------------------------------------
    def reset(self, positions, envs: Optional[Sequence[int]] = None) -> None:
        """Fresh map + one initial fire at (x, y) per listed env (simulation.py:202-214)."""
        xy = np.ascontiguousarray(np.asarray(positions, dtype=np.int32).reshape(-1, 2))
        n = xy.shape[0]
        if envs is None:
            ev = None
        else:
            ev = np.ascontiguousarray(np.asarray(envs, dtype=np.int32).reshape(-1))
            if ev.shape[0] != n:
                raise ValueError("reset: one (x, y) per env expected")
        _lib.check(self._lib.sfb_reset(self._h, _ptr(ev) if ev is not None else None, n, _ptr(xy)))

    def apply_points(self, points) -> None:
        """(env, x, y, BurnStatus) rows -> fire_map[env][y, x] = kind (mitigation.py:60-80)."""
        pts = np.ascontiguousarray(np.asarray(points, dtype=np.int32).reshape(-1, 4))
        if pts.shape[0]:
            _lib.check(self._lib.sfb_apply_points(self._h, _ptr(pts), pts.shape[0]))

    def set_fire_map(self, maps: np.ndarray, env0: int = 0) -> None:
        m = np.ascontiguousarray(np.asarray(maps).astype(np.int8, copy=False)).reshape(-1, self.H, self.W)
        _lib.check(self._lib.sfb_set_fire_map(self._h, int(env0), m.shape[0], _ptr(m)))

    # -- the hot path ---------------------------------------------------------------------
    def step(self, n: int = 1, sync: bool = True) -> None:
        _lib.check(self._lib.sfb_step(self._h, int(n), 1 if sync else 0))

    def step_timed(self, n: int = 1) -> float:
        """n steps bracketed by CUDA events on the engine's stream; returns milliseconds."""
        ms = C.c_float()
        _lib.check(self._lib.sfb_step_timed(self._h, int(n), C.byref(ms)))
        return float(ms.value)

    def update(self, maps: np.ndarray, env0: int = 0) -> np.ndarray:
        """
        `manager.update(fire_map)` with host buffers: `maps` (int8, [n, H, W] or [H, W],
        C-contiguous) is uploaded, stepped once and overwritten in place; returns the
        GameStatus of each env.
        """
        if maps.dtype != np.int8 or not maps.flags.c_contiguous:
            raise ValueError("update: maps must be a C-contiguous int8 array")
        n = maps.size // (self.H * self.W)
        status = np.empty(n, dtype=np.int32)
        _lib.check(self._lib.sfb_update(self._h, int(env0), n, _ptr(maps), _ptr(status)))
        return status

    def constant_spread_update(self, maps: np.ndarray, rate_of_spread: int, env0: int = 0) -> None:
        """`fire_map = ConstantSpreadFireManager.update(fire_map)` (fire.py:754-787) with host maps, in place."""
        if maps.dtype != np.int8 or not maps.flags.c_contiguous:
            raise ValueError("constant_spread_update: maps must be a C-contiguous int8 array")
        n = maps.size // (self.H * self.W)
        _lib.check(self._lib.sfb_constant_spread_update(self._h, int(env0), n, _ptr(maps), int(rate_of_spread)))

    def synchronize(self) -> None:
        _lib.check(self._lib.sfb_synchronize(self._h))

    # -- results --------------------------------------------------------------------------
    def fire_map(self, env0: int = 0, n: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        n = self.E - env0 if n is None else n
        if out is None:
            out = np.empty((n, self.H, self.W), dtype=np.int8)
        elif out.dtype != np.int8 or not out.flags.c_contiguous or out.size != n * self.H * self.W:
            raise ValueError("fire_map: out must be a C-contiguous int8 array of n*H*W cells")
        _lib.check(self._lib.sfb_get_fire_map(self._h, int(env0), int(n), _ptr(out)))
        return out

    def sync_fire_maps(self, mirror: np.ndarray) -> int:
        """
        Bring the caller's host mirror (int8 [E, H, W], C-contiguous, the same array on every
        call, not modified in between) up to date.  With `track_changes=True` only the cells
        that changed since the previous call cross PCIe.  Returns the number of patched cells
        (-1: everything was downloaded).
        """
        if mirror.dtype != np.int8 or not mirror.flags.c_contiguous or mirror.size != self.E * self.H * self.W:
            raise ValueError("sync_fire_maps: mirror must be a C-contiguous int8 array of E*H*W cells")
        n = C.c_int64()
        _lib.check(self._lib.sfb_sync_fire_maps(self._h, _ptr(mirror), C.byref(n)))
        return int(n.value)

    def set_tracking(self, enabled: bool) -> None:
        """Pause / resume the change log (engines created with track_changes=True)."""
        _lib.check(self._lib.sfb_set_tracking(self._h, 1 if enabled else 0))

    def plane(self, which: str, env: int = 0) -> np.ndarray:
        pid, dt = {"burn": (_lib.PLANE_BURN, np.float64), "ros": (_lib.PLANE_ROS, np.float64),
                   "age": (_lib.PLANE_AGE, np.int32), "status": (_lib.PLANE_STATUS, np.int8),
                   "ignition": (_lib.PLANE_IGNITION, np.int32)}[which]  # fmt: skip
        out = np.empty((self.H, self.W), dtype=dt)
        _lib.check(self._lib.sfb_get_plane(self._h, int(env), pid, _ptr(out)))
        return out

    def status(self):
        """(GameStatus int32[E], elapsed_time float64[E], update() calls int32[E])."""
        st = np.empty(self.E, dtype=np.int32)
        el = np.empty(self.E, dtype=np.float64)
        n = np.empty(self.E, dtype=np.int32)
        _lib.check(self._lib.sfb_get_status(self._h, _ptr(st), _ptr(el), _ptr(n)))
        return st, el, n

    def fire_map_device(self):
        """Zero-copy int8 [E, H, W] BurnStatus view in device memory, as a torch tensor."""
        import torch

        p = C.c_void_p()
        _lib.check(self._lib.sfb_fire_map_device(self._h, C.byref(p)))

        class _Iface:
            __cuda_array_interface__ = {
                "shape": (self.E, self.H, self.W), "typestr": "|i1", "data": (p.value, False),
                "version": 3, "strides": None,
            }  # fmt: skip

        return torch.as_tensor(_Iface(), device=f"cuda:{self.device}")

    def static_device(self):
        """Zero-copy float32 CUDA tensor [n_sets, H, W, 8] over the static records (w_0, delta, M_x, sigma,
        U, U_dir, slope_mag, slope_dir per cell; n_sets = 1 when the envs share one terrain).  A strided
        view of the library's own memory: read-only for the caller."""
        import torch

        p, plane, pitch, sets = C.c_void_p(), C.c_int64(), C.c_int32(), C.c_int32()
        _lib.check(self._lib.sfb_static_device(self._h, C.byref(p), C.byref(plane), C.byref(pitch), C.byref(sets)))

        class _Iface:
            __cuda_array_interface__ = {
                "shape": (int(sets.value), self.H, self.W, 8), "typestr": "<f4", "data": (p.value, False), "version": 3,
                "strides": (int(plane.value) * 32, int(pitch.value) * 32, 32, 4),
            }  # fmt: skip

        return torch.as_tensor(_Iface(), device=f"cuda:{self.device}")

    # -- slab mode (one grid split in row slabs across engines / GPUs) ---------------------
    def state_device(self):
        """(device pointer, plane cells, pitch cells, bytes per cell) of the packed state plane."""
        ptr, plane, pitch, cb = C.c_void_p(), C.c_int64(), C.c_int32(), C.c_int32()
        _lib.check(self._lib.sfb_state_device(self._h, C.byref(ptr), C.byref(plane), C.byref(pitch), C.byref(cb)))
        return int(ptr.value), int(plane.value), int(pitch.value), int(cb.value)

    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        _lib.check(self._lib.sfb_ipc_export(self._h, buf))
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        _lib.check(self._lib.sfb_ipc_open(self.device, C.create_string_buffer(handle, 64), C.byref(ptr)))
        return int(ptr.value)

    def ipc_close(self, ptr: int) -> None:
        _lib.check(self._lib.sfb_ipc_close(self.device, C.c_void_p(ptr)))

    def set_halo(self, top_row: int, top_plane: int, bottom_row: int, bottom_plane: int) -> None:
        _lib.check(self._lib.sfb_set_halo(self._h, C.c_void_p(top_row or None), int(top_plane),
                                          C.c_void_p(bottom_row or None), int(bottom_plane)))  # fmt: skip

    def slab_mailbox(self):
        """(device pointer of this slab's mailbox, its byte offset from the state plane)."""
        ptr, off = C.c_void_p(), C.c_int64()
        _lib.check(self._lib.sfb_slab_mailbox(self._h, C.byref(ptr), C.byref(off)))
        return int(ptr.value), int(off.value)

    def slab_connect(self, rank: int, world: int, peer_mailboxes) -> None:
        arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_mailboxes])
        _lib.check(self._lib.sfb_slab_connect(self._h, int(rank), int(world), arr))

    def step_slab(self, n: int = 1) -> None:
        """n steps coordinated with the other slabs through peer memory; asynchronous."""
        _lib.check(self._lib.sfb_step_slab(self._h, int(n)))

    def step_sweep(self) -> None:
        _lib.check(self._lib.sfb_step_sweep(self._h))

    def step_eval(self) -> None:
        _lib.check(self._lib.sfb_step_eval(self._h))

    def flags_tensors(self):
        """Two int32 CUDA tensors [E, 8] viewing the double-buffered per-env records; index
        them with `self.flags_parity()`.  MAX-reducing the one in flight across slabs ORs the
        any_live / any_cand flags."""
        import torch

        p, n = C.c_void_p(), C.c_int64()
        _lib.check(self._lib.sfb_flags_device(self._h, C.byref(p), C.byref(n)))
        cur = int(p.value)
        nbytes = int(n.value) * 4
        base = cur - nbytes if self._parity_probe() else cur
        out = []
        for k in range(2):
            class _Iface:
                __cuda_array_interface__ = {"shape": (self.E, 8), "typestr": "<i4", "data": (base + k * nbytes, False),
                                            "version": 3, "strides": None}  # fmt: skip
            out.append(torch.as_tensor(_Iface(), device=f"cuda:{self.device}"))
        return out

    def _parity_probe(self) -> int:
        par = C.c_int32()
        _lib.check(self._lib.sfb_get_parity(self._h, C.byref(par)))
        return int(par.value)

    def flags_parity(self) -> int:
        return self._parity_probe()

    def set_stream(self, stream: int) -> None:
        _lib.check(self._lib.sfb_set_stream(self._h, C.c_void_p(stream or None)))

    # -- introspection --------------------------------------------------------------------
    @property
    def stream(self) -> int:
        p = C.c_void_p()
        _lib.check(self._lib.sfb_get_stream(self._h, C.byref(p)))
        return int(p.value or 0)

    def launch_counts(self):
        a, b = C.c_int64(), C.c_int64()
        _lib.check(self._lib.sfb_get_launch_counts(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_kernel_timing(self, enabled: bool) -> None:
        _lib.check(self._lib.sfb_set_kernel_timing(self._h, 1 if enabled else 0))

    def kernel_ms(self):
        """(k_sweep ms, k_rows ms, k_eval ms, steps) accumulated since timing was enabled; a bitboard handle has no
        first kernel (0) and reports k_tiles as the second."""
        a, r, b, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        _lib.check(self._lib.sfb_get_kernel_ms(self._h, C.byref(a), C.byref(r), C.byref(b), C.byref(n)))
        return float(a.value), float(r.value), float(b.value), int(n.value)

    def row_tasks(self):
        a, b = C.c_int64(), C.c_int64()
        _lib.check(self._lib.sfb_get_row_tasks(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def unit_stats(self):
        """(units the last step listed, units of the handle); equal without unit skipping."""
        a, b, m = C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(self._lib.sfb_get_unit_stats(self._h, C.byref(a), C.byref(b), C.byref(m)))
        return int(a.value), int(b.value)

    def unit_mode(self) -> str:
        """'bits' (the bitboard step: k_tiles + k_eval on self-maintained tile lists; the default for big handles with
        max_fire_duration <= 7), 'dense' (every unit swept), 'chunks' (flagged chunks swept), 'rows' (flagged rows are
        the row tasks) or 'lists' (the list-driven step: no units, one watch list)."""
        a, b, m = C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(self._lib.sfb_get_unit_stats(self._h, C.byref(a), C.byref(b), C.byref(m)))
        return ("dense", "chunks", "rows", "lists", "bits")[int(m.value)]

    def queue_stats(self):
        a, b, o = C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(self._lib.sfb_get_queue_stats(self._h, C.byref(a), C.byref(b), C.byref(o)))
        return int(a.value), int(b.value), bool(o.value)

    def front_stats(self) -> dict:
        """List handles: what k_front did since the previous call (the counters are reset).  Bitboard handles: what
        k_tiles did while kernel timing was on (examined = cells of the tiles looked at, entries_read = tiles,
        neighbourhoods_read = control-line cells left to k_eval)."""
        a = (C.c_int64 * 7)()
        _lib.check(self._lib.sfb_get_front_stats(self._h, a, 7))
        names = ("examined", "candidates", "ignited", "pruned", "joined", "entries_read", "neighbourhoods_read")
        return {k: int(v) for k, v in zip(names, a)}

    def debug_stall(self, microseconds: int) -> None:
        """Test knob: keep the engine's stream busy for a while before the next call's work."""
        _lib.check(self._lib.sfb_debug_stall(self._h, int(microseconds)))

    def device_bytes(self) -> int:
        b = C.c_int64()
        _lib.check(self._lib.sfb_device_bytes(self._h, C.byref(b)))
        return int(b.value)


def rate_of_spread(direction, w_0, delta, M_x, sigma, U, U_dir, slope_mag, slope_dir, *,
                   h=8000.0, S_T=0.0555, S_e=0.01, p_p=32.0, M_f=0.03, device: int = 0) -> np.ndarray:  # fmt: skip
    """
    Device evaluation of `compute_rate_of_spread` (simfire/world/rothermel.py:4-136) for n
    pairs; `direction[i]` indexes the neighbour order of fire.py:211-221, the other arrays
    are the destination cell's values.  Returns float64 ft/min.
    """
    lib = _lib.load()
    d = np.ascontiguousarray(np.asarray(direction).astype(np.int8))
    n = d.shape[0]
    rec = np.empty((n, 8), dtype=np.float32)
    for i, a in enumerate((w_0, delta, M_x, sigma, U, U_dir, slope_mag, slope_dir)):
        rec[:, i] = np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)).astype(np.float32)
    part = np.array([h, S_T, S_e, p_p, M_f], dtype=np.float32)
    out = np.empty(n, dtype=np.float64)
    _lib.check(lib.sfb_rate_of_spread(int(device), _ptr(d), _ptr(rec), _ptr(part), n, _ptr(out)))
    return out
