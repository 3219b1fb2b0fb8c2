import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import torch
from bench import make_workload, bench_starts
from simfire_b200 import FireEngine
wl, E, shared = make_workload("target")
eng = FireEngine(wl.H, wl.W, E, shared_static=True, track_changes=True, **wl.engine_kwargs())
eng.set_static(wl.planes); starts = bench_starts(wl, E, 0); eng.reset(starts)
eng.set_tracking(False); eng.step(65); eng.set_tracking(True)
maps = torch.empty((E, wl.H, wl.W), dtype=torch.int8, pin_memory=True).numpy()
pts = torch.empty((E, 4), dtype=torch.int32, pin_memory=True).numpy()
rng = np.random.default_rng(5)
pts[:,0]=np.arange(E); pts[:,3]=3
eng.sync_fire_maps(maps)
def step():
    pts[:,1]=rng.integers(0,wl.W,E); pts[:,2]=rng.integers(0,wl.H,E)
    eng.apply_points(pts); eng.step(1, sync=False); return eng.sync_fire_maps(maps)
for _ in range(3): step()
t0=time.perf_counter(); n=0
for _ in range(30): n+=step()
dt=(time.perf_counter()-t0)/30
print(f"threads={os.environ.get('SFB_HOST_THREADS')} e2e {dt*1e3:.3f} ms/step, {n/30:.0f} changes/step", flush=True)
