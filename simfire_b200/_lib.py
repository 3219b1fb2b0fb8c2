"""ctypes binding of libsimfire_b200.so (include/simfire_b200.h).  No fallback: if the
library is missing or the call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB

ABI_VERSION = 1

# sfb_flags
DIAGONAL_SPREAD, ATTENUATE_LINE_ROS, SHARED_STATIC, KEEP_ROS, HAS_MAX_TIME, WIDE_CELLS, SWEEP_LDG = 1, 2, 4, 8, 16, 32, 64
TRACK_CHANGES = 128
KEEP_IGNITION = 256
UNIT_SKIP_OFF, UNIT_SKIP_ON, UNIT_CHUNKS, STEP_GRAPH, FRONT_LISTS = 512, 1024, 2048, 4096, 8192
FRONT_BITS = 16384
# sfb_state_plane
PLANE_BURN, PLANE_ROS, PLANE_AGE, PLANE_STATUS, PLANE_IGNITION = 0, 1, 2, 3, 4
STATIC_PLANES = ("w_0", "delta", "M_x", "sigma", "U", "U_dir", "slope_mag", "slope_dir")

# every symbol include/simfire_b200.h declares
EXPORTS = (
    "sfb_create", "sfb_destroy", "sfb_last_error", "sfb_abi_version", "sfb_set_static",
    "sfb_set_static_all", "sfb_reset", "sfb_apply_points", "sfb_set_fire_map", "sfb_step",
    "sfb_step_timed", "sfb_update", "sfb_synchronize", "sfb_get_fire_map", "sfb_get_plane",
    "sfb_get_status", "sfb_fire_map_device", "sfb_get_stream", "sfb_get_launch_counts",
    "sfb_set_kernel_timing", "sfb_get_kernel_ms", "sfb_get_queue_stats", "sfb_device_bytes",
    "sfb_rate_of_spread", "sfb_sync_fire_maps", "sfb_state_device", "sfb_ipc_export", "sfb_ipc_open",
    "sfb_ipc_close", "sfb_set_halo", "sfb_step_sweep", "sfb_step_eval", "sfb_flags_device", "sfb_set_stream",
    "sfb_slab_mailbox", "sfb_slab_connect", "sfb_step_slab", "sfb_set_tracking",
    "sfb_set_elevation", "sfb_get_row_tasks", "sfb_get_unit_stats", "sfb_debug_stall", "sfb_get_front_stats", "sfb_get_parity", "sfb_constant_spread_update", "sfb_static_device",
)  # fmt: skip


class SfbParams(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("E", C.c_int32), ("max_fire_duration", C.c_int32), ("flags", C.c_int32),
        ("rows_per_chunk", C.c_int32), ("pixel_scale", C.c_double), ("update_rate", C.c_double),
        ("max_time", C.c_double), ("h", C.c_float), ("S_T", C.c_float), ("S_e", C.c_float),
        ("p_p", C.c_float), ("M_f", C.c_float), ("env_groups", C.c_int32),
        ("queue_capacity", C.c_int64), ("slab_y0", C.c_int32), ("slab_total_H", C.c_int32),
    ]  # fmt: skip


class SfbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libsimfire_b200 error {code}: {message}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load the shared library (built by simfire_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise ImportError(
            f"{LIB} is missing: build it with `python -m simfire_b200.build` "
            "(simfire_b200 has no CPU or PyTorch fallback for the fire-spread step)"
        )
    lib = C.CDLL(LIB)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    proto = {
        "sfb_create": (C.c_int, [C.POINTER(SfbParams), C.POINTER(vp)]),
        "sfb_destroy": (None, [vp]),
        "sfb_last_error": (C.c_char_p, []),
        "sfb_abi_version": (C.c_int, []),
        "sfb_set_static": (C.c_int, [vp, i32, i32, vp]),
        "sfb_set_static_all": (C.c_int, [vp, i32, vp]),
        "sfb_reset": (C.c_int, [vp, vp, i32, vp]),
        "sfb_apply_points": (C.c_int, [vp, vp, i64]),
        "sfb_set_fire_map": (C.c_int, [vp, i32, i32, vp]),
        "sfb_step": (C.c_int, [vp, i32, i32]),
        "sfb_step_timed": (C.c_int, [vp, i32, C.POINTER(C.c_float)]),
        "sfb_update": (C.c_int, [vp, i32, i32, vp, vp]),
        "sfb_synchronize": (C.c_int, [vp]),
        "sfb_get_fire_map": (C.c_int, [vp, i32, i32, vp]),
        "sfb_get_plane": (C.c_int, [vp, i32, i32, vp]),
        "sfb_sync_fire_maps": (C.c_int, [vp, vp, C.POINTER(i64)]),
        "sfb_state_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]),
        "sfb_ipc_export": (C.c_int, [vp, vp]),
        "sfb_ipc_open": (C.c_int, [i32, vp, C.POINTER(vp)]),
        "sfb_ipc_close": (C.c_int, [i32, vp]),
        "sfb_set_halo": (C.c_int, [vp, vp, i64, vp, i64]),
        "sfb_step_sweep": (C.c_int, [vp]),
        "sfb_step_eval": (C.c_int, [vp]),
        "sfb_flags_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(i64)]),
        "sfb_set_stream": (C.c_int, [vp, vp]),
        "sfb_slab_mailbox": (C.c_int, [vp, C.POINTER(vp), C.POINTER(i64)]),
        "sfb_slab_connect": (C.c_int, [vp, i32, i32, C.POINTER(vp)]),
        "sfb_step_slab": (C.c_int, [vp, i32]),
        "sfb_set_tracking": (C.c_int, [vp, i32]),
        "sfb_set_elevation": (C.c_int, [vp, i32, vp]),
        "sfb_get_status": (C.c_int, [vp, vp, vp, vp]),
        "sfb_fire_map_device": (C.c_int, [vp, C.POINTER(vp)]),
        "sfb_get_stream": (C.c_int, [vp, C.POINTER(vp)]),
        "sfb_get_launch_counts": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64)]),
        "sfb_set_kernel_timing": (C.c_int, [vp, i32]),
        "sfb_get_kernel_ms": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                        C.POINTER(i64)]),
        "sfb_get_row_tasks": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64)]),
        "sfb_get_unit_stats": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]),
        "sfb_get_queue_stats": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]),
        "sfb_device_bytes": (C.c_int, [vp, C.POINTER(i64)]),
        "sfb_debug_stall": (C.c_int, [vp, i32]),
        "sfb_static_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]),
        "sfb_constant_spread_update": (C.c_int, [vp, i32, i32, vp, i32]),
        "sfb_get_parity": (C.c_int, [vp, C.POINTER(i32)]),
        "sfb_get_front_stats": (C.c_int, [vp, C.POINTER(i64), i32]),
        "sfb_rate_of_spread": (C.c_int, [i32, vp, vp, vp, i64, vp]),
    }
    for name, (res, args) in proto.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.sfb_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB}: ABI version {lib.sfb_abi_version()}, binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise SfbError(rc, load().sfb_last_error().decode("utf-8", "replace"))
