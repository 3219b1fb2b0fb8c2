"""
CPU-side check of the CUDA sources themselves: tests/emu/ compiles simfire_b200/csrc/sfb.cu
(C ABI, host glue and every kernel, unchanged) with g++ against a small fiber emulator of
the CUDA execution model (warps, collectives, persistent blocks, TMA boxes, graph replay) and
the `-m gpu` parity tests are re-run against that build in a subprocess.

This is test infrastructure only.  The emulator library is built into tests/emu/_build/
(git- and gpurun-ignored), is never loaded by the package on its own (a test has to point
SFB_LIB at it) and does not exist on the GPU box, where the same tests run against the real
library.  It checks kernel and host logic (work lists, env groups, change logs, layouts),
not timing, memory ordering or races.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# need a CUDA context of torch's own (device tensors around raw pointers): not emulated
NEEDS_TORCH_CUDA = "not observation_tensor and not row_slabs and not device_view"


def _run(args, timeout=900):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    from build_emu import build

    lib = build()
    env = dict(os.environ, SFB_LIB=lib, SFB_EMULATED="1")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", *args]
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = "\n".join((res.stdout + res.stderr).splitlines()[-25:])
    assert res.returncode == 0, f"emulated run failed:\n{tail}"
    return res.stdout


def test_emulator_library_is_not_the_product():
    """The emulator build carries a marker symbol the product library must not have, and lives
    outside the package."""
    import ctypes

    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    from build_emu import OUT, build

    lib = ctypes.CDLL(build())
    assert lib.sfb_emu_marker() == 1
    assert not os.path.commonpath([OUT, os.path.join(ROOT, "simfire_b200")]) == os.path.join(ROOT, "simfire_b200")
    product = os.path.join(ROOT, "simfire_b200", "libsimfire_b200.so")
    if os.path.exists(product):
        assert not hasattr(ctypes.CDLL(product), "sfb_emu_marker")
    ignore = open(os.path.join(ROOT, ".gpurunignore")).read()
    assert "tests/emu/_build" in ignore  # never travels to the GPU box


def test_gpu_parity_suite_under_emulation():
    out = _run(["tests/test_gpu_parity.py", "-k", NEEDS_TORCH_CUDA])
    assert " passed" in out


def test_gpu_api_suite_under_emulation():
    out = _run(["tests/test_gpu_api.py", "tests/test_spread_graph.py", "-k", NEEDS_TORCH_CUDA])
    assert " passed" in out


def test_small_scale_cases_under_emulation():
    out = _run(["tests/test_gpu_scale.py", "-k",
                "degenerate_and_ragged or fire_duration_limits or cell_life_cycle"])  # fmt: skip
    assert " passed" in out
