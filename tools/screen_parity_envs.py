"""Dev-container tool: which envs of a bench batch can be compared bit for bit with the CPU reference?

    python tools/screen_parity_envs.py target 160 24

fire_map parity is exact only while no ignition test `burn > pixel_scale` (fire.py:568) is decided by
less than the float32 libm differences between NumPy and the device (~1e-7 relative on a burn value).
For every env of the batch (ignition cells of bench.bench_starts) this steps the NumPy oracle on the
window the fire cannot leave and records the smallest relative distance |burn - ps| / ps over all
ignition tests; envs whose margin stays above 2e-5 (the admission rule of tests/golden/gen_golden.py)
go into bench.PARITY_ENVS and tests/test_gpu_scale.py.  Deterministic: same seeds, same list."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import bench_starts, make_workload  # noqa: E402
from oracle.dense_numpy import DenseFire, DenseParams  # noqa: E402
from oracle.reference_runner import window_around  # noqa: E402


def margin_of(wl, start, n):
    y0, x0, h, w = window_around(start, n, wl.H, wl.W)
    planes = {k: np.ascontiguousarray(np.broadcast_to(v, (wl.H, wl.W))[y0 : y0 + h, x0 : x0 + w]) for k, v in wl.planes.items()}
    sim = DenseFire(planes, DenseParams(**wl.engine_kwargs()), (int(start[0]) - x0, int(start[1]) - y0))
    m = np.inf
    for _ in range(n):
        before = sim.burn.copy()
        if sim.step() != 1:
            break
        ch = sim.burn != before
        if ch.any():
            m = min(m, float(np.min(np.abs(sim.burn[ch] - wl.pixel_scale) / wl.pixel_scale)))
    return m, int((sim.status != 0).sum()), sim.step_count


if __name__ == "__main__":
    name, n, count = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    wl, E, _ = make_workload(name)
    starts = bench_starts(wl, E, 0)
    good = []
    for e in range(count):
        m, cells, steps = margin_of(wl, starts[e], n)
        print(f"env {e} start {tuple(starts[e])}: margin {m:.2e}, {cells} cells touched, {steps} updates", flush=True)
        if m > 2e-5 and steps == n:
            good.append(e)
    print("good:", good)
