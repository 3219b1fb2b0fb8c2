"""TEST-ONLY: runs bench.py's GPU arm end to end against the emulator build (tiny workload) so
that a Python-level mistake in the bench shows up in the CPU suite and not on the GPU box.
torch.cuda is stubbed just enough for that; the numbers printed mean nothing."""
import os
import runpy
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *_a, **_k: None
torch.cuda.synchronize = lambda *_a, **_k: None
_empty = torch.empty


def _empty_unpinned(*a, **k):
    k.pop("pin_memory", None)
    return _empty(*a, **k)


torch.empty = _empty_unpinned
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
