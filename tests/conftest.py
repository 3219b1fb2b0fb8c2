import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible and -m gpu was not forced.  The parity
    tests at the BASELINE sizes (tests/test_gpu_scale.py) run FIRST: the driver uses `-x`, and an unrelated
    failure in an earlier file must not keep the benchmarked configuration from being checked."""
    items.sort(key=lambda it: 0 if "test_gpu_scale" in it.nodeid else 1)  # stable: everything else keeps its order
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu or os.environ.get("SFB_EMULATED") == "1":
        return  # (SFB_EMULATED: tests/test_emu_parity.py re-runs GPU tests against the fiber emulator build)
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
