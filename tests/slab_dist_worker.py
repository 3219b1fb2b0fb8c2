"""torchrun worker for tests/test_gpu_multigpu.py: one slab per rank over NVLink peer memory
must reproduce a single-engine run of the whole grid (computed on rank 0)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simfire_b200 import FireEngine  # noqa: E402
from simfire_b200.sharding import RankContext  # noqa: E402
from simfire_b200.slab import SlabGrid  # noqa: E402
from simfire_b200.workloads import synthetic_operational  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ctx = RankContext.from_env(backend="nccl", device_id=torch.device("cuda", local))
    H, W, E = 203, 300, 2
    wl = synthetic_operational(H, W, seed=8, patch=8)
    kw = dict(wl.engine_kwargs(), attenuate_line_ros=True, max_time=90.0)
    mid = H // ctx.world
    starts = [(150, mid - 1), (40, mid + 2)]
    lines = [(e, x, y, 3 + (x % 3)) for e in range(E) for y in (20, mid, 150) for x in range(10, 290, 2)]
    grid = SlabGrid(H, W, wl.planes, ctx=ctx, E=E, device=local, sync=os.environ.get("SFB_SLAB_SYNC", "p2p"), **kw)
    grid.reset(starts)
    grid.apply_points(lines)
    ref = None
    if ctx.rank == 0:
        ref = FireEngine(H, W, E, device=local, sweep_ldg=True, **kw)
        ref.set_static(wl.planes)
        ref.reset(starts)
        ref.apply_points(lines)
    ok = True
    for it in range(25):
        n = 1 + it % 4
        grid.step(n)
        fm = grid.fire_map()
        if ctx.rank == 0:
            ref.step(n)
            same = np.array_equal(fm, ref.fire_map())
            st_ok = all(np.array_equal(a, b) for a, b in zip(grid.status(), ref.status()))
            if not (same and st_ok):
                ok = False
                print(f"MISMATCH at block {it}: maps {same} status {st_ok}", flush=True)
                break
    ms = grid.step_timed(20)
    if ctx.rank == 0:
        burned = int((fm == 2).sum())
        print(f"SLAB_DIST {'OK' if ok and burned > 500 else 'FAIL'} world={ctx.world} burned={burned} "
              f"ms_per_step={ms / 20:.4f}", flush=True)
        ref.close()
    grid.close()
    ctx.close()


if __name__ == "__main__":
    main()
