"""CPU-side checks of the C-ABI boundary: the library builds, loads, exports every symbol
include/simfire_b200.h declares, and refuses to compute without a GPU (no fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from simfire_b200.build import build_library

    return build_library()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "simfire_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfb_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree(lib_path):
    from simfire_b200 import _lib

    decl = declared_symbols()
    assert len(decl) >= 20
    assert sorted(_lib.EXPORTS) == decl


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.sfb_abi_version() == 1


def test_params_struct_layout_matches_header(lib_path):
    from simfire_b200 import _lib

    # 8 x int32, 3 x double, 5 x float + int32, int64, 2 x int32 -> 96 bytes, no padding
    assert ctypes.sizeof(_lib.SfbParams) == 96
    assert _lib.SfbParams.pixel_scale.offset == 32
    assert _lib.SfbParams.queue_capacity.offset == 80


def test_flag_constants_match_the_header():
    """The ctypes binding's copies of `enum sfb_flags` / plane ids are the header's values."""
    from simfire_b200 import _lib

    src = open(os.path.join(ROOT, "include", "simfire_b200.h")).read()
    enum = dict((k, int(v)) for k, v in re.findall(r"\b(SFB_[A-Z_0-9]+)\s*=\s*(\d+)\s*[,/}]", src))
    want = {"SFB_DIAGONAL_SPREAD": _lib.DIAGONAL_SPREAD, "SFB_ATTENUATE_LINE_ROS": _lib.ATTENUATE_LINE_ROS,
            "SFB_SHARED_STATIC": _lib.SHARED_STATIC, "SFB_KEEP_ROS": _lib.KEEP_ROS, "SFB_HAS_MAX_TIME": _lib.HAS_MAX_TIME,
            "SFB_WIDE_CELLS": _lib.WIDE_CELLS, "SFB_SWEEP_LDG": _lib.SWEEP_LDG, "SFB_TRACK_CHANGES": _lib.TRACK_CHANGES,
            "SFB_KEEP_IGNITION": _lib.KEEP_IGNITION, "SFB_UNIT_SKIP_OFF": _lib.UNIT_SKIP_OFF, "SFB_UNIT_SKIP_ON": _lib.UNIT_SKIP_ON,
            "SFB_UNIT_CHUNKS": _lib.UNIT_CHUNKS, "SFB_STEP_GRAPH": _lib.STEP_GRAPH, "SFB_FRONT_LISTS": _lib.FRONT_LISTS,
            "SFB_FRONT_BITS": _lib.FRONT_BITS}  # fmt: skip
    for name, value in want.items():
        assert enum.get(name) == value, (name, enum.get(name), value)
    flags = [v for k, v in enum.items() if k in want]
    assert len(set(flags)) == len(flags) and all(v & (v - 1) == 0 for v in flags)  # distinct single bits


def test_no_cpu_fallback(lib_path):
    """Without a CUDA device the product path must fail loudly."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from simfire_b200 import FireEngine, SfbError

    with pytest.raises(SfbError, match="no CUDA device"):
        FireEngine(16, 16, 1, pixel_scale=50.0, update_rate=1.0, max_fire_duration=4)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under simfire_b200/ may import or call it."""
    pkg = os.path.join(ROOT, "simfire_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "reference" not in [m for m in re.findall(r"sys\.path\.\w+\(.*?/root/(\w+)", text)], f


def test_host_compiled_rothermel_matches_golden(tmp_path):
    """The device function's source, compiled for the host (test-only), must reproduce the
    reference's operation order: exact zeros pattern, median error at the ulp level."""
    import subprocess

    import numpy as np

    so = tmp_path / "host_rothermel.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                    os.path.join(ROOT, "tests", "host_rothermel.cpp")], check=True)  # fmt: skip
    lib = ctypes.CDLL(str(so))
    z = np.load(os.path.join(ROOT, "tests", "golden", "rothermel_pairs.npz"))
    n = len(z["direction"])
    rec = np.ascontiguousarray(np.stack([z[k] for k in ("w_0", "delta", "M_x", "sigma", "U", "U_dir", "slope_mag", "slope_dir")], 1).astype(np.float32))
    d = np.ascontiguousarray(z["direction"].astype(np.int8))
    part = z["consts"].astype(np.float32)
    out = np.zeros(n)
    vp = ctypes.c_void_p
    lib.host_rate_of_spread(vp(d.ctypes.data), vp(rec.ctypes.data), vp(part.ctypes.data), ctypes.c_longlong(n), vp(out.ctypes.data))
    R = z["R"]
    assert np.array_equal(out == 0, R == 0)
    nz = R != 0
    rel = np.abs(out[nz] - R[nz]) / np.abs(R[nz])
    assert np.median(rel) < 5e-7
    assert np.quantile(rel, 0.99) < 1e-5
