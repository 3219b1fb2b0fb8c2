"""Tests that need more than one GPU on the box (skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("sync", ["p2p", "nccl"])
def test_row_slabs_across_gpus_match_single_grid(sync):
    """cfg5-style domain decomposition: one slab per rank, halo rows read from peer memory
    (CUDA IPC over NVLink), flags all-reduced with NCCL; result equals a single-GPU run."""
    n = min(_n_gpus(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "slab_dist_worker.py")]  # fmt: skip
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, SFB_SLAB_SYNC=sync))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "SLAB_DIST OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
