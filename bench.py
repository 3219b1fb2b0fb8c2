#!/usr/bin/env python
"""
bench.py -- cell-updates/s of the fire-spread stepper (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one `RothermelFireManager.update` (simfire/game/managers/fire.py:616-719) over
every env of the batch.  Default workload = the configuration the metric is quoted on
(north_star target): 2048 x 2048 synthetic flat terrain x 1024 independent envs per GPU
sharing one terrain (weak scaling: each rank owns 1024 envs, no collective on the data
path).  One JSON line is printed by rank 0.

value     device-timed (CUDA events on the engine's stream), state resident in HBM
e2e       same metric through the public batched API with HOST buffers: per step the
          mitigation points go host -> device from pinned memory and every env's fire_map
          comes back device -> host into pinned memory
roofline  dominant kernel (k_sweep) against the measured HBM peak
cpu_baseline / --impl reference
          the NumPy oracle port of the reference algorithm on the host cores (the reference
          is pure Python and does not travel to the GPU box)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cell_updates_per_s"
UNIT = "cell-updates/s"

# bytes per cell-update of a kernel that streams every plane each step
SURVEY_BYTES_PER_ENV_STATIC = 52.0
SURVEY_BYTES_SHARED = lambda E: 20.0 + 32.0 / E  # noqa: E731


def workload_spec(name: str):
    """name -> (H, W, envs per GPU, shared_static, flat, description)"""
    specs = {
        # north_star target: "2048^2 x 1024-env batch at 1 GPU", synthetic flat terrain
        "target": (2048, 2048, 1024, True, True),
        # BASELINE configs[2]/[3]: 512^2 x 1024 envs per GPU, per-env terrain replaced by shared
        "cfg3": (512, 512, 1024, True, False),
        # the same with a terrain of its own for every env (16 distinct terrains, env i gets i % 16)
        "cfg3_perenv": (512, 512, 1024, False, False),
        # BASELINE configs[1]
        "cfg2": (1024, 1024, 1, False, False),
        "small": (256, 256, 64, True, True),
        # BASELINE configs[4]: one 8192^2 grid, row slabs across the GPUs (strong scaling)
        "cfg5": (8192, 8192, 1, False, False),
    }
    return specs[name]


def make_workload(name: str):
    from simfire_b200.workloads import synthetic_operational

    H, W, E, shared, flat = workload_spec(name)
    wl = synthetic_operational(H, W, seed=0, flat=flat)
    return wl, E, shared


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    REASONS = {
        0x0000000000000004: "sw_power_cap",
        0x0000000000000008: "hw_slowdown",
        0x0000000000000020: "sw_thermal_slowdown",
        0x0000000000000040: "hw_thermal_slowdown",
        0x0000000000000080: "hw_power_brake_slowdown",
    }

    def __init__(self, index: int, period_s: float = 0.002):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _run(self):
        nv = self.nvml
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                for bit, nm in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nvml is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}  # fmt: skip


# --------------------------------------------------------------------------------------
# CPU legs: the UNMODIFIED reference (oracle/_ref staged by oracle/build_ref.py, driven through
# oracle/reference_runner.py) on the host cores; the NumPy port (oracle/dense_numpy.py) only when
# the reference was not staged.  Test/bench infrastructure: never on the product path.
# --------------------------------------------------------------------------------------
def bench_config(args, wl, E, shared, world):
    """The workload description both arms print (identical dicts for identical arguments)."""
    return {"workload": args.workload, "grid": [wl.H, wl.W], "envs_per_gpu": E, "envs_total": E * world,
            "static_planes": "shared" if shared else "per-env", "terrain": wl.description,
            "burn_in_steps": args.burn_in, "parallelism": f"env-sharded x{world}, no collective",
            "l2": "no flush between steps: the state advances every step, and what a step touches (the cells of the "
                  "moving fire fronts inside planes far larger than the 126 MB L2) is what a production rollout "
                  "touches; measured per-step footprint in timing_notes"}  # fmt: skip


def bench_starts(wl, E, rank):
    """Ignition cell of every env of a rank (the GPU arm and the CPU arm light the same fires)."""
    return wl.burnable_starts(E, seed=1000 + rank)


def _cpu_sim(wl, start, window=None, kind=None):
    """One env on the host: (simulation, kind).  kind 'reference' = RothermelFireManager.update itself."""
    from oracle import reference_runner as rr

    if kind is None:
        kind = "reference" if rr.available() else "port"
    if kind == "reference":
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return rr.from_workload(wl, start, window=window), kind
    from oracle.dense_numpy import DenseFire, DenseParams

    y0, x0, h, w = window if window is not None else (0, 0, wl.H, wl.W)
    planes = {k: np.ascontiguousarray(np.broadcast_to(v, (wl.H, wl.W))[y0 : y0 + h, x0 : x0 + w]) for k, v in wl.planes.items()}
    sim = DenseFire(planes, DenseParams(**wl.engine_kwargs()), (int(start[0]) - x0, int(start[1]) - y0))
    sim.init_seconds = 0.0
    return sim, kind


# envs of the `target` / `cfg3` batches (rank 0) whose ignition tests all keep a relative distance
# > 2e-5 from the threshold for 160 updates (tools/screen_parity_envs.py): the device's float32
# pow / exp / cos differ from NumPy's by an ulp or two, which moves burn values by ~1e-7 relative
PARITY_ENVS = {"target": [0, 2, 3, 4, 6, 7, 8, 9], "cfg3": [1, 3, 4, 7, 8, 10, 11, 13]}


def cpu_baseline_leg(args, wl, E, budget_s: float = 12.0, gpu_maps=None):
    """One host core steps ONE env of the bench batch (env 0: same terrain, same ignition cell, full
    grid) through the burn-in and warm-up and is timed on the updates that follow -- the first of
    the updates the GPU arm times.  `gpu_maps(starts, n)` returns the device's fire_maps of envs lit
    at `starts` after n updates: parity is stated in the same run, against the reference itself --
    the timed env on the full grid, and PARITY_ENVS on windows their fires provably cannot leave."""
    from oracle import reference_runner as rr

    starts = bench_starts(wl, E, 0)
    sim, kind = _cpu_sim(wl, starts[0])
    pre = args.burn_in + args.warmup
    t0 = time.perf_counter()
    for _ in range(pre):
        sim.step()
    pre_s = time.perf_counter() - t0
    n, t0 = 0, time.perf_counter()
    while True:
        sim.step()
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= max(1, args.steps):
            break
    what = ("RothermelFireManager.update of the unmodified reference (oracle/_ref, fire.py:616)" if kind == "reference"
            else "oracle/dense_numpy.py (NumPy port: the reference is not staged under oracle/_ref)")
    out = {"value": wl.H * wl.W * n / dt, "unit": UNIT, "cores": 1, "kind": kind,
           "sample": f"{what}, env 0 of the batch on the full {wl.H}x{wl.W} grid, updates {pre + 1}..{pre + n} "
                     f"in {dt:.1f} s after {pre} untimed updates ({pre_s:.1f} s)",
           "init_s": round(float(sim.init_seconds), 2),
           "init_note": "manager construction, not in `value`: one networkx node per pixel (fire.py:380)"}  # fmt: skip
    if gpu_maps is None:
        return out
    try:  # a checker: it must never cost the run its line
        par = {"against": kind, "full_grid": None, "windows": None}
        got = gpu_maps(starts[:1], pre + n)[0]
        par["full_grid"] = {"env": 0, "updates": pre + n, "fire_map_equal": bool(np.array_equal(got, sim.status)),
                            "cells_burning_or_burned": int((sim.status == 1).sum() + (sim.status == 2).sum())}  # fmt: skip
        envs = [e for e in PARITY_ENVS.get(args.workload, list(range(min(E, 4)))) if e < E]
        n_par = args.parity_updates
        if envs and n_par > 0:
            maps = gpu_maps(starts[envs], n_par)
            equal, cells = [], 0
            for k, e in enumerate(envs):
                win = rr.window_around(starts[e], n_par, wl.H, wl.W)
                ws, _ = _cpu_sim(wl, starts[e], window=win, kind=kind)
                for _ in range(n_par):
                    ws.step()
                y0, x0, h, w = win
                inside = np.array_equal(maps[k][y0 : y0 + h, x0 : x0 + w], ws.status)
                outside = int((maps[k] != 0).sum()) == int((maps[k][y0 : y0 + h, x0 : x0 + w] != 0).sum())
                equal.append(bool(inside and outside))
                cells += int((ws.status == 1).sum() + (ws.status == 2).sum())
            par["windows"] = {"envs": envs, "updates": n_par, "fire_map_equal": equal, "all_equal": all(equal),
                              "cells_burning_or_burned": cells,
                              "note": "each env re-run on the host inside the window its fire cannot leave in that many "
                                      "updates (oracle.reference_runner.window_around); the device map must equal it "
                                      "inside and be untouched outside"}  # fmt: skip
        out["parity"] = par
    except Exception as exc:  # pragma: no cover
        out["parity"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


def _ref_worker(job):
    name, env, start, pre, n_steps, window, kind = job
    os.environ["OMP_NUM_THREADS"] = "1"
    wl, _, _ = make_workload(name)
    sim, kind = _cpu_sim(wl, start, window=window, kind=kind)
    for _ in range(pre):
        sim.step()
    t = []
    for _ in range(n_steps):
        t0 = time.perf_counter()
        sim.step()
        t.append(time.perf_counter() - t0)
    return t, float(sim.init_seconds), int((sim.status != 0).sum())


def reference_arm(args):
    """`--impl reference`: the reference's own `RothermelFireManager.update` on every host core -- one
    process per core, each stepping one env of the bench batch (same terrain, same ignition cells,
    same burn-in, full grid), i.e. a bounded sample (as many envs as cores) of the GPU arm's batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    from oracle import reference_runner as rr

    kind = "reference" if rr.available() else "port"
    wl, E, shared = make_workload(args.workload)
    if args.envs:
        E = args.envs
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64, E))
    try:  # a reference process holds ~0.55 KB per cell (object arrays + the networkx graph)
        import psutil

        per_proc = (600.0 if kind == "reference" else 120.0) * wl.H * wl.W + 4e8
        workers = max(1, min(workers, int(0.7 * psutil.virtual_memory().available / per_proc)))
    except Exception:
        pass
    pre = args.burn_in + args.warmup
    # bounded: ~40 ns per cell and update for the reference's whole-grid object-array passes
    window_note = "full grid"
    windows = [None] * workers
    est = (pre + args.steps) * wl.H * wl.W * (60e-9 if kind == "reference" else 300e-9)
    starts = bench_starts(wl, E, 0)
    if est > 300.0:
        windows = [rr.window_around(starts[i], pre + args.steps, wl.H, wl.W) for i in range(workers)]
        window_note = "each env on the window its fire cannot leave (the full grid would take ~%.0f s)" % est
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        t_start = time.perf_counter()
        res = pool.map(_ref_worker, [(args.workload, i, tuple(int(v) for v in starts[i]), pre, args.steps, windows[i], kind)
                                     for i in range(workers)])  # fmt: skip
        wall_all = time.perf_counter() - t_start
    # like the GPU arm: the job's time is the slowest worker's (max over "ranks")
    secs = max(float(np.sum(r[0])) for r in res)
    value = workers * wl.H * wl.W * args.steps / secs
    what = ("RothermelFireManager.update of the unmodified reference (oracle/_ref, fire.py:616)" if kind == "reference"
            else "oracle/dense_numpy.py (NumPy port: the reference is not staged under oracle/_ref)")
    sample = (f"{what}: {workers} processes (of {cores} host cores) x 1 env each = envs 0..{workers - 1} of the {E}-env batch, "
              f"{window_note}, updates {pre + 1}..{pre + args.steps} timed after {pre} untimed ones; manager construction "
              f"{np.mean([r[1] for r in res]):.1f} s per process (not timed); pool wall {wall_all:.1f} s")  # fmt: skip
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8+f32/f64", "data": "synthetic",
        "config": bench_config(args, wl, E, shared, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def load_traffic_note(workload: str, kernel: str, row_tasks=None, work_items=None):
    """Per-launch DRAM bytes of `kernel` from the committed ncu captures, if they match the workload:
    (bytes, note).  profiles/ncu_sweep_summary.json: the dense sweep (same launch shape as any dense run);
    profiles/ncu_kernel_traffic.json: the front-proportional kernels, captured at the task / item counts the
    file names and scaled linearly to this run's counts.  A label, not a measurement of the timed run."""
    try:
        if kernel.startswith("k_sweep"):
            with open(os.path.join(ROOT, "profiles", "ncu_sweep_summary.json")) as f:
                j = json.load(f)
            if j.get("workload") == workload:
                return j.get("dram_bytes_per_launch"), "ncu --set full, same launch shape"
        else:
            with open(os.path.join(ROOT, "profiles", "ncu_kernel_traffic.json")) as f:
                j = json.load(f)
            if j.get("workload") == workload and kernel in j.get("kernels", {}):
                b = j["kernels"][kernel]["dram_bytes_per_launch"]
                scale = 1.0
                if kernel == "k_rows" and row_tasks:
                    scale = row_tasks / max(1, j.get("row_tasks_at_capture", row_tasks))
                if kernel == "k_eval" and work_items:
                    scale = work_items / max(1, j.get("work_items_at_capture", work_items))
                return (b * scale,
                        f"ncu --set full at {j.get('row_tasks_at_capture')} row tasks / {j.get('work_items_at_capture')} work items "
                        f"per launch ({b / 1e6:.1f} MB), scaled x{scale:.2f} to this run's counts")
    except Exception:
        pass
    return None, None


def gpu_arm_slab(args):
    """cfg5: a single 8192 x 8192 grid, one row slab per GPU, halo rows over NVLink peer memory."""
    import torch

    from simfire_b200.sharding import RankContext
    from simfire_b200.slab import SlabGrid

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ctx = RankContext.from_env(backend="nccl", device_id=torch.device("cuda", local))
    wl, E, _ = make_workload(args.workload)
    H, W = wl.H, wl.W
    grid = SlabGrid(H, W, wl.planes, ctx=ctx, n_slabs=1, E=E, device=local, track_changes=not args.no_track,
                    sync=args.slab_sync, **wl.engine_kwargs())  # fmt: skip
    grid.reset([wl.init_pos])
    grid.step(args.burn_in)
    grid.step(args.warmup)
    eng = grid.engines[0]
    l0 = eng.launch_counts()[1]
    with ClockSampler(local) as clocks:
        torch.cuda.synchronize()
        ctx.barrier()
        ms = grid.step_timed(args.steps)
        torch.cuda.synchronize()
        ctx.barrier()
    launches = eng.launch_counts()[1] - l0
    # diagnostic (--diag-back-to-back): the same K steps once more right away, the GPU still busy, no barrier before
    ms_b2b = ctx.max(eng.step_timed(args.steps)) if args.diag_back_to_back else None
    ms_max = ctx.max(ms)
    value = H * W * E * args.steps / (ms_max * 1e-3)
    # e2e: one mitigation point in, this rank's rows of the fire_map mirrored on the host, per step
    rows = grid.slabs[ctx.rank][1] if ctx.world > 1 else H
    mirror = torch.empty((E, rows, W), dtype=torch.int8, pin_memory=True).numpy()
    rng = np.random.default_rng(5)
    eng.sync_fire_maps(mirror)
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    torch.cuda.synchronize()
    ctx.barrier()
    t0 = time.perf_counter()
    changes = 0
    for _ in range(e2e_steps):
        grid.apply_points([(0, int(rng.integers(0, W)), int(rng.integers(0, H)), 3)])
        grid.step(1)
        changes += max(0, eng.sync_fire_maps(mirror))
    torch.cuda.synchronize()
    ctx.barrier()
    e2e_value = H * W * E * e2e_steps / ctx.max(time.perf_counter() - t0)
    st, el, nst = grid.status()
    burned = ctx.sum(float((mirror == 2).sum()))
    if ctx.rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8+f32/f64", "data": "synthetic",
            "config": {"workload": args.workload, "grid": [H, W], "envs_total": E, "terrain": wl.description,
                       "burn_in_steps": args.burn_in,
                       "parallelism": (f"{ctx.world} row slabs, halo rows read from peer memory (CUDA IPC over NVLink), "
                                       + ("per-step flags and hand-shakes stored into the peers' mailboxes (no NCCL)"
                                          if args.slab_sync == "p2p" else "2 NCCL all-reduces of <= 32 B per step"))
                                      if ctx.world > 1 else "1 GPU",
                       "l2": "a 67 MB state plane fits the 126 MB L2: this workload is launch/latency-bound"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 16,
                    "d2h_bytes_per_step": int(8 * changes / e2e_steps) + 12, "steps": e2e_steps,
                    "api": "SlabGrid.apply_points + step + sync_fire_maps (host mirror of this rank's rows)"},
            "gpu_launches": int(launches),
            "sanity": {"running": int(st[0]), "burned_cells": int(burned), "steps_done": int(nst[0])},
        }  # fmt: skip
        print(json.dumps(line), flush=True)
    grid.close()
    ctx.close()


def full_burn_arm(args):
    """`--full-burn` ("also report a full-burn run for cfg1-2"): one env from its
    configured ignition cell until the reference's loop would stop (GameStatus.QUIT), launched in
    blocks of 64 updates; device-timed.  cfg1 is also replayed on the NumPy oracle and compared."""
    import torch

    from simfire_b200 import FireEngine
    from simfire_b200.workloads import cfg1_functional_flat, synthetic_operational

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the stepper has no CPU path")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.cuda.set_device(local)
    wl = cfg1_functional_flat(128) if args.workload == "cfg1" else synthetic_operational(1024, 1024, seed=0)
    with FireEngine(wl.H, wl.W, 1, device=local, **wl.engine_kwargs()) as eng:
        eng.set_static(wl.planes)
        eng.reset([wl.init_pos])
        eng.step(3)  # warm-up launches (the fire is three updates old when the clock starts)
        ms, updates = 0.0, 3
        with ClockSampler(local) as clocks:
            while True:
                ms += eng.step_timed(64)
                st, el, n = eng.status()
                updates = int(n[0])
                if not st[0] or updates > 200000:
                    break
        final = eng.fire_map(0, 1)[0]
        line = {"metric": METRIC, "value": wl.H * wl.W * (updates - 3) / (ms * 1e-3), "unit": UNIT, "n_gpus": 1,
                "steps": updates - 3, "warmup": 3, "ms_per_step": ms / max(1, updates - 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8+f32/f64", "data": "synthetic",
                "config": {"workload": args.workload + " full burn", "grid": [wl.H, wl.W], "envs_total": 1,
                           "terrain": wl.description, "start": list(wl.init_pos),
                           "note": "updates until GameStatus.QUIT, enqueued 64 at a time; an env that has quit is "
                                   "skipped by the kernels, so the last block is partly idle"},
                "clocks": clocks.summary(), "gpu_launches": int(eng.launch_counts()[1]),
                "result": {"updates_until_quit": updates, "elapsed_time_min": float(el[0]),
                           "burned_cells": int((final == 2).sum()), "unburned_cells": int((final == 0).sum())}}  # fmt: skip
    if args.workload == "cfg1":
        sim, kind = _cpu_sim(wl, wl.init_pos)
        t0 = time.perf_counter()
        while sim.step() == 1:
            pass
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": wl.H * wl.W * sim.step_count / dt, "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"{'the unmodified reference (oracle/_ref)' if kind == 'reference' else 'oracle/dense_numpy.py'}, "
                                          f"the same full burn: {sim.step_count} updates in {dt:.1f} s",
                                "parity": {"fire_map_equal": bool(np.array_equal(final, sim.status)),
                                           "updates_equal": sim.step_count == updates}}  # fmt: skip
    print(json.dumps(line), flush=True)


FRONT_KW = {"auto": {}, "lists": dict(front_lists=True), "bits": dict(front_bits=True), "rows": dict(unit_skip=True),
            "chunks": dict(unit_skip=True, unit_chunks=True), "dense": dict(unit_skip=False)}


def load_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f).get("hbm_gbs", 6650.0)), "measured"
    except Exception:
        return 6650.0, "fallback"


def roofline_lists(eng, args, cells_rank, peak_gbs, peak_src, workload):
    """The list-driven step: k_front is the step (k_tail only exists with attenuation).  Algorithmic
    bytes per launch, counted from the kernel's own counters over the per-launch-timed pass
    (DESIGN.md section 4): per list entry read 8 B and, if it stays, 8 B written; per examined cell
    its state byte; per cell that looks at its neighbourhood 8 more state bytes; per candidate the
    8-byte rate of its (cell, direction) pair and the float64 burn value read and written; one byte
    per ignition / burn-out; a 4-byte bitmap word read and written per cell that joins the list."""
    eng.front_stats()  # reset the counters
    eng.set_kernel_timing(True)
    eng.step(args.roofline_steps)
    front_ms, ros_ms, tail_ms, n_t = eng.kernel_ms()
    eng.set_kernel_timing(False)
    fs = {k: v / n_t for k, v in eng.front_stats().items()}
    entries, cap, went_dense = eng.queue_stats()
    kept = entries  # the list the next step reads = what the last step kept
    front_bytes = (fs["entries_read"] * 8.0 + kept * 8.0 + fs["examined"] * 1.0 + fs["neighbourhoods_read"] * 8.0 +
                   fs["candidates"] * 24.0 + fs["ignited"] + fs["pruned"] + fs["joined"] * 8.0)
    front_s = front_ms / n_t * 1e-3
    tail_s = tail_ms / n_t * 1e-3
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_kernel_traffic.json")) as f:
            j = json.load(f)
        if j.get("workload") == workload and "k_front" in j.get("kernels", {}):
            k = j["kernels"]["k_front"]
            # scaled from the capture's list length to this run's (the traffic is proportional to the entries)
            traffic = k["dram_bytes_per_launch"] * fs["entries_read"] / max(1.0, k["entries_read_at_capture"])
            traffic_src = ("committed ncu capture (profiles/ncu_kernel_traffic.json: %.1f MB at %d entries), scaled to this "
                           "run's %d entries per launch; not a measurement of the timed run" %
                           (k["dram_bytes_per_launch"] / 1e6, k["entries_read_at_capture"], fs["entries_read"]))
    except Exception:
        pass
    achieved = front_bytes / front_s / 1e9
    return {
        "bound": "latency", "kernel": "k_front", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
        "frac": achieved / peak_gbs, "peak_source": f"{peak_src} copy bandwidth (MEASURED_PEAKS.json hbm_gbs)",
        "traffic": traffic, "traffic_source": traffic_src,
        "note": "front-proportional gather kernel: one thread per watched cell, scattered byte reads of the 3x3 "
                "neighbourhood out of L2, one 8-byte table read and a float64 read-modify-write per candidate.  It moves "
                "a few tens of MB per launch, so HBM bandwidth is not what bounds it (memory latency and the launch "
                "are); the HBM-bound kernel of this design is the dense TMA sweep (`dense_sweep`).",
        "ms_per_launch": front_s * 1e3, "bytes_per_launch": front_bytes, "bytes_per_cell_update": front_bytes / cells_rank,
        "share_of_step": front_s / max(1e-12, front_s + tail_s),
        "per_launch": {k: round(v, 1) for k, v in fs.items()}, "list_entries": entries, "list_capacity": cap,
        "went_dense": went_dense,
        "kernel_ms_per_launch": {"k_front": front_s * 1e3, "k_tail": tail_s * 1e3},
        "front": "lists",
    }


def roofline_bits(eng, args, cells_rank, peak_gbs, peak_src, workload, ring, step_ms, launches_per_step):
    """The bitboard front end: k_tiles + k_eval, k_tiles being the step.  Algorithmic bytes
    per launch of k_tiles, from the kernel's own counters over the per-launch-timed
    pass: per tile (16 rows x 30 columns) its 8-byte task, the 16 words of the ignitable and control-line
    planes and of the expiring sprite plane, 18 words of each of the ring - 1 source planes (16 rows + the row
    above and below); per candidate the 8-byte rate of its (cell, direction) pair and the float64 burn value read
    and written; one state byte per ignition / burn-out; per tile up to three 128-byte plane rows written
    back and one flag byte.
    Launch duration: a step is these two kernels back to back on one stream, so the time of the dominant one
    is taken as (the device-timed step of the timed region) x (its share of the two in a pass with CUDA events
    around every launch); the event-bracketed times themselves carry ~10 us of launch and event latency per
    launch, which is not kernel time, and are reported next to it."""
    eng.front_stats()  # reset the counters
    eng.set_kernel_timing(True)
    eng.step(args.roofline_steps)
    _, tiles_ms, eval_ms, n_t = eng.kernel_ms()
    eng.set_kernel_timing(False)
    fs = {k: v / n_t for k, v in eng.front_stats().items()}
    tiles, cand = fs["entries_read"], fs["candidates"]
    units_listed, units_total = eng.unit_stats()
    q_entries, q_cap, q_ovf = eng.queue_stats()
    tiles_s, eval_s = tiles_ms / n_t * 1e-3, eval_ms / n_t * 1e-3
    tile_rows = round(fs["examined"] / max(1.0, tiles) / 30.0)  # 16 (two tiles per warp at a time) or 32
    tiles_bytes = (tiles * (8.0 + 3 * 4.0 * tile_rows + (ring - 1) * (tile_rows + 2) * 4.0 + 2.0 + 8.0) + cand * 24.0 +
                   fs["ignited"] * (1.0 + 12.0) + fs["pruned"] * (1.0 + 4.0))
    eval_bytes = q_entries * (8.0 + 16.0) + 64.0 * eng.E
    bracketed_ms = {"k_tiles": tiles_s * 1e3, "k_eval": eval_s * 1e3}
    if launches_per_step < 1.5:  # no control lines anywhere: k_tiles closes the step itself, there is no k_eval launch
        bracketed_ms = {"k_tiles": tiles_s * 1e3}
        eval_s = 0.0
    share = {k: v / (tiles_s + eval_s) / 1e3 for k, v in bracketed_ms.items()}
    kernel_ms = {k: step_ms * share[k] for k in bracketed_ms}
    kernel_bytes = {"k_tiles": tiles_bytes + (eval_bytes if "k_eval" not in bracketed_ms else 0.0), "k_eval": eval_bytes}
    dominant = max(kernel_ms, key=kernel_ms.get)
    dom_s = kernel_ms[dominant] * 1e-3
    achieved = kernel_bytes[dominant] / dom_s / 1e9
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_kernel_traffic.json")) as f:
            j = json.load(f)
        if j.get("workload") == workload and dominant == "k_tiles" and "k_tiles" in j.get("kernels", {}):
            k = j["kernels"]["k_tiles"]
            traffic = k["dram_bytes_per_launch"] * tiles / max(1.0, k["tiles_at_capture"])
            traffic_src = ("committed ncu capture, not a measurement of the timed run: %.1f MB read + %.1f MB written at %d tiles per "
                           "launch with a cold L2 for every replay pass (profiles/r02/r02_k_tiles_ncu.json), scaled to this run's %d tiles"
                           % (k["dram_bytes_read"] / 1e6, k["dram_bytes_write"] / 1e6, k["tiles_at_capture"], tiles))
    except Exception:
        pass
    return {
        "bound": "latency", "kernel": dominant, "env_groups_timed_one_after_the_other": True, "achieved": achieved,
        "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
        "peak_source": f"{peak_src} copy bandwidth (MEASURED_PEAKS.json hbm_gbs)",
        "traffic": traffic, "traffic_source": traffic_src,
        "note": "front-proportional kernels: a step touches a few tens of MB (bit planes of the listed tiles, rate / burn "
                "of the candidates), so HBM bandwidth is not what bounds it -- dependent-load latency and the launches "
                "are; the HBM-bound kernel of this design is the dense TMA sweep (`dense_sweep`).",
        "kernels": {k: {"ms_per_launch": kernel_ms[k], "bytes_per_launch": kernel_bytes[k],
                        "achieved": kernel_bytes[k] / (kernel_ms[k] * 1e-3) / 1e9 if kernel_ms[k] > 0 else None}
                    for k in kernel_ms},
        "unit_skipping": {"on": True, "mode": "bits", "units_listed": units_listed, "units_total": units_total,
                          "cells_swept_per_step": 0.0, "cells_per_step": cells_rank},
        "kernel_ms_per_launch": kernel_ms, "kernel_ms_per_launch_event_bracketed": bracketed_ms,
        "launch_duration_method": "device-timed step of the timed region x the kernel's share of the event-bracketed pass",
        "launches_per_step": launches_per_step,
        "bytes_per_launch": kernel_bytes[dominant],
        "bytes_per_cell_update": kernel_bytes[dominant] / cells_rank, "ms_per_launch": dom_s * 1e3,
        "share_of_step": share[dominant],
        "per_launch": {k: round(v, 1) for k, v in fs.items()},
        "tile_rows": tile_rows, "row_tasks_per_step": tiles, "work_items_per_step": cand, "items_left_to_k_eval": q_entries,
        "queue_overflowed": q_ovf, "front": "bits",
    }  # fmt: skip


def roofline_sweeps(eng, args, cells_rank, peak_gbs, peak_src, workload):
    """The sweep front ends (--front rows | chunks | dense): front end + k_rows + k_eval."""
    eng.set_kernel_timing(True)
    eng.step(args.roofline_steps)
    sweep_ms, rows_ms, eval_ms, n_t = eng.kernel_ms()
    eng.set_kernel_timing(False)
    q_entries, q_cap, q_ovf = eng.queue_stats()
    row_tasks, _ = eng.row_tasks()
    units_listed, units_total = eng.unit_stats()  # of the same (last) step of the pass
    sweep_s, rows_s, eval_s = sweep_ms / n_t * 1e-3, rows_ms / n_t * 1e-3, eval_ms / n_t * 1e-3
    unit_mode = eng.unit_mode()
    skipping = unit_mode != "dense"
    if unit_mode == "rows":  # nothing is swept: k_row_list reads one flag byte per (env, row, strip)
        sweep_cells = 0.0
        sweep_bytes = units_total * 1.0 + row_tasks * 8.0
        front_kernel = "k_row_list"
    else:  # 1 B per cell of every listed unit (+ an 8-byte row task per warp-row that needs a look)
        sweep_cells = cells_rank * (units_listed / max(1, units_total))
        sweep_bytes = sweep_cells * 1.0 + row_tasks * 8.0 + (units_total + 8.0 * units_listed if skipping else 0.0)
        front_kernel = "k_sweep_" + args.sweep
    # k_rows: 8-byte task + three 512-byte rows per task, 8 B per work item; k_eval: item, 48-byte record, burn r/w
    rows_bytes = row_tasks * (3 * 512 + 8.0) + q_entries * 8.0
    eval_bytes = q_entries * (8.0 + 48.0 + 8.0 + 8.0)
    kernel_ms = {front_kernel: sweep_s * 1e3, "k_rows": rows_s * 1e3, "k_eval": eval_s * 1e3}
    kernel_bytes = {front_kernel: sweep_bytes, "k_rows": rows_bytes, "k_eval": eval_bytes}
    dominant = max(kernel_ms, key=kernel_ms.get)
    dom_s = kernel_ms[dominant] * 1e-3
    achieved = kernel_bytes[dominant] / dom_s / 1e9
    traffic, traffic_note = load_traffic_note(workload, dominant, row_tasks, q_entries)
    if dominant.startswith("k_sweep") and skipping:
        traffic, traffic_note = None, None  # the committed capture is of the dense sweep
    bound = {"k_rows": "issue", "k_eval": "latency"}.get(dominant, "hbm")
    note = {"k_rows": "issue-bound, not HBM-bound: ~76 % of the issue slots busy, DRAM at 12-14 % (profiles/r01b_kernels.json)",
            "k_eval": "latency-bound gather/scatter on a work queue (DRAM ~30 %, issue slots ~38 %)"}.get(
                dominant, "streaming kernel: the roofline that matters is HBM bandwidth")
    return {
        "bound": bound, "kernel": dominant, "env_groups_timed_one_after_the_other": True, "achieved": achieved,
        "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
        "peak_source": f"{peak_src} copy bandwidth (MEASURED_PEAKS.json hbm_gbs)",
        "traffic": traffic, "traffic_source": ("committed ncu capture, not a measurement of the timed run: " + traffic_note) if traffic_note else None,
        "note": note,
        "kernels": {k: {"ms_per_launch": kernel_ms[k], "bytes_per_launch": kernel_bytes[k],
                        "achieved": kernel_bytes[k] / (kernel_ms[k] * 1e-3) / 1e9 if kernel_ms[k] > 0 else None}
                    for k in kernel_ms},
        "unit_skipping": {"on": skipping, "mode": unit_mode, "units_listed": units_listed, "units_total": units_total,
                          "cells_swept_per_step": sweep_cells, "cells_per_step": cells_rank},
        "kernel_ms_per_launch": kernel_ms, "bytes_per_launch": kernel_bytes[dominant],
        "bytes_per_cell_update": kernel_bytes[dominant] / cells_rank, "ms_per_launch": dom_s * 1e3,
        "share_of_step": dom_s / (sweep_s + rows_s + eval_s),
        "row_tasks_per_step": row_tasks, "work_items_per_step": q_entries, "queue_overflowed": q_ovf, "front": unit_mode,
    }  # fmt: skip


def gpu_arm(args):
    import torch

    from simfire_b200 import FireEngine
    from simfire_b200.sharding import RankContext

    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the stepper has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    ctx = RankContext.from_env(backend="nccl", device_id=torch.device("cuda", local))
    rank, world = ctx.rank, ctx.world
    # host threads that patch the fire_map mirror: share the box's cores between the ranks
    os.environ.setdefault("SFB_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // max(1, world))))

    wl, E, shared = make_workload(args.workload)
    if args.envs:
        E = args.envs
    H, W = wl.H, wl.W
    front_kw = dict(FRONT_KW[args.front])
    if args.front != "lists":
        front_kw.update(rows_per_chunk=args.rows_per_chunk, sweep_ldg=(args.sweep == "ldg"), env_groups=args.env_groups)

    def make_engine(**kw):
        return FireEngine(H, W, E, shared_static=shared, device=local, **kw, **wl.engine_kwargs())

    eng = make_engine(track_changes=not args.no_track, **front_kw)
    if args.workload == "cfg3_perenv":
        from simfire_b200.workloads import synthetic_operational

        variants = [synthetic_operational(H, W, seed=k) for k in range(16)]
        for e in range(E):
            eng.set_static(variants[e % 16].planes, env=e)
        starts = np.stack([variants[e % 16].burnable_starts(1, seed=1000 + rank * E + e)[0] for e in range(E)])
    else:
        eng.set_static(wl.planes)
        starts = bench_starts(wl, E, rank)
    eng.reset(starts)
    eng.step(args.burn_in)  # untimed: let the fronts develop so the timed steps see real fires

    host_barrier = os.environ.get("SFB_BENCH_BARRIER", "host") == "host"

    def barrier():
        torch.cuda.synchronize()
        ctx.barrier(host=host_barrier)  # NCCL still carries the reductions of the timings (ctx.max / ctx.sum)
        torch.cuda.synchronize()

    # device-resident phases (value, roofline): nothing leaves the GPU, so the change log that
    # feeds the host mirror is paused; it is resumed for the end-to-end phase
    if not args.no_track:
        eng.set_tracking(False)
    eng.step(args.warmup)
    l0 = eng.launch_counts()[1]
    with ClockSampler(local) as clocks:
        barrier()
        ms = eng.step_timed(args.steps)
        barrier()
    launches = eng.launch_counts()[1] - l0
    # diagnostic (--diag-back-to-back): the same K steps once more right away, the GPU still busy, no barrier before
    ms_b2b = ctx.max(eng.step_timed(args.steps)) if args.diag_back_to_back else None
    ms_max = ctx.max(ms)
    ms_min = -ctx.max(-ms)  # the fastest rank: ranks light different fires, so their steps are not equally long
    cells_rank = H * W * E
    cells_per_step = cells_rank * world
    value = cells_per_step * args.steps / (ms_max * 1e-3)

    # ---- dominant-kernel timing for the roofline (separate pass, every launch bracketed by events)
    peak_gbs, peak_src = load_peak()
    unit_mode = eng.unit_mode()
    if unit_mode == "bits":
        roof = roofline_bits(eng, args, cells_rank, peak_gbs, peak_src, args.workload, wl.max_fire_duration + 1, ms_max / args.steps, launches / args.steps)
    else:
        roof = (roofline_lists if unit_mode == "lists" else roofline_sweeps)(eng, args, cells_rank, peak_gbs, peak_src, args.workload)

    if not args.no_track:
        eng.set_tracking(True)

    # ---- end to end through the batched API with host buffers
    if args.mirror == "thp":
        # the mirror is only written by host threads (patches) and by one initial download: ordinary
        # memory on transparent huge pages keeps the scattered patch writes out of the page walker
        import mmap

        mirror_mm = mmap.mmap(-1, E * H * W, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        if hasattr(mmap, "MADV_HUGEPAGE"):
            mirror_mm.madvise(mmap.MADV_HUGEPAGE)
        maps_np = np.frombuffer(mirror_mm, dtype=np.int8).reshape(E, H, W)
        maps_np[...] = 0  # touch every page before the timed region
    else:
        pinned_maps = torch.empty((E, H, W), dtype=torch.int8, pin_memory=True)
        maps_np = pinned_maps.numpy()
    pinned_pts = torch.empty((E, 4), dtype=torch.int32, pin_memory=True)
    pts_np = pinned_pts.numpy()
    rng = np.random.default_rng(5 + rank)
    e2e_steps = max(3, min(args.steps, args.e2e_steps))

    # the caller's actions: one control-line point per env per step, (env, x, y, kind) rows
    n_extra = 8
    actions = np.empty((e2e_steps + 2 + n_extra, E, 4), dtype=np.int32)
    actions[:, :, 0] = np.arange(E)
    actions[:, :, 1] = rng.integers(0, W, actions.shape[:2])
    actions[:, :, 2] = rng.integers(0, H, actions.shape[:2])
    actions[:, :, 3] = 3  # BurnStatus.FIRELINE
    e2e_calls = [0.0, 0.0, 0.0]

    def e2e_step(i, timed=False):
        ta = time.perf_counter()
        pts_np[...] = actions[i]
        eng.apply_points(pts_np)          # host -> device
        tb = time.perf_counter()
        eng.step(1, sync=False)
        tc = time.perf_counter()
        n = eng.sync_fire_maps(maps_np)   # device -> host: changed cells only, patched into the mirror
        if timed:
            td = time.perf_counter()
            e2e_calls[0] += tb - ta
            e2e_calls[1] += tc - tb
            e2e_calls[2] += td - tc
        return n

    e2e_step(0)  # first call downloads every map once
    e2e_step(1)
    barrier()
    t0 = time.perf_counter()
    e2e_changes = 0
    for i in range(e2e_steps):
        e2e_changes += max(0, e2e_step(2 + i))
    barrier()
    e2e_s = time.perf_counter() - t0
    for i in range(n_extra):  # where a step's wall time goes, call by call (outside the timed region)
        e2e_step(2 + e2e_steps + i, timed=True)
    e2e_calls = [round(1e3 * v / n_extra, 4) for v in e2e_calls]
    # the patched mirror must equal a full download (checked outside the timed region)
    e2e_ok = bool(np.array_equal(maps_np[: min(E, 16)], eng.fire_map(0, min(E, 16))))
    e2e_value = cells_per_step * e2e_steps / ctx.max(e2e_s)

    # ---- the same loop for a consumer that keeps the observation on the device (an RL policy on the GPU
    # reads fire_map_device / the packed state): mitigation points in, per-env GameStatus and elapsed_time out
    if not args.no_track:
        eng.set_tracking(False)  # nobody drains the change log in this loop: such a consumer does not ask for it
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        pts_np[...] = actions[i]
        eng.apply_points(pts_np)
        eng.step(1, sync=False)
        eng.status()  # device -> host: E x (status, elapsed, update count); synchronises
    barrier()
    status_only_value = cells_per_step * e2e_steps / ctx.max(time.perf_counter() - t0)
    if not args.no_track:
        eng.set_tracking(True)

    # ---- sanity: the timed steps really advanced fires
    st, el, nsteps = eng.status()
    burned = int((maps_np[: min(E, 8)] == 2).sum())
    running = int(st.sum())

    # ---- the same batch from ignition on: how the step time moves with the age of the fires
    age = None
    if args.age_curve > 0 and args.workload != "cfg3_perenv":
        if not args.no_track:
            eng.set_tracking(False)
        eng.reset(starts)
        marks = [m for m in (50, 100, 200, 300, 500, 750, 1000, 1500, 2000) if m <= args.age_curve]
        blocks, done, total_ms = [], 0, 0.0
        for m in marks:
            blk_ms = ctx.max(eng.step_timed(m - done))
            blocks.append({"updates": [done + 1, m], "ms_per_step": blk_ms / (m - done),
                           "value": cells_per_step * (m - done) / (blk_ms * 1e-3)})
            total_ms += blk_ms
            done = m
        st2 = eng.status()[0]
        age = {"blocks": blocks, "updates": done, "ms_per_step": total_ms / done,
               "value": cells_per_step * done / (total_ms * 1e-3), "envs_still_running": int(st2.sum()),
               "note": "every env re-lit at its ignition cell, then stepped; `value` above is the K timed steps after "
                       "the burn-in, this is ignition -> update %d (dense-counted the same way)" % done}

    if rank != 0:
        ctx.close()
        return

    survey_b = SURVEY_BYTES_SHARED(E) if shared else SURVEY_BYTES_PER_ENV_STATIC
    step_s = ms_max / args.steps * 1e-3
    roof["survey_model"] = {"bytes_per_cell_update": survey_b, "achieved": cells_rank * survey_b / step_s / 1e9,
                            "frac": cells_rank * survey_b / step_s / 1e9 / peak_gbs,
                            "note": "SURVEY.md 8d counts a kernel that streams every plane each step; this design touches "
                                    "only the cells at the fire fronts, so the figure is not a roofline fraction"}
    log_b = 8
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8+f32/f64", "data": "synthetic",
        "config": bench_config(args, wl, E, shared, world),
        "timing_notes": {"front": unit_mode, "timed_updates": [args.burn_in + args.warmup + 1, args.burn_in + args.warmup + args.steps],
                         "footprint_per_step_MB": roof["bytes_per_launch"] / 1e6,
                         "ms_per_step_slowest_rank": ms_max / args.steps, "ms_per_step_fastest_rank": ms_min / args.steps,
                         **({"ms_per_step_next_K_steps_without_a_barrier_before": ms_b2b / args.steps} if ms_b2b else {}),
                         "barrier": ("host-side (gloo) barrier + torch.cuda.synchronize() on both sides of the timed region; an NCCL "
                                     "barrier kernel in front of it costs every rank ~45 us before its first step kernel starts "
                                     "(profiles/r02/r03g_*: rank 0, same fires, 0.0274 ms/step after a host barrier, 0.0297 after an "
                                     "NCCL one), which a 0.56 ms region reads as an 8 % scaling loss; NCCL carries the reductions")
                                    if host_barrier else "NCCL barrier + torch.cuda.synchronize() on both sides of the timed region",
                         "ranks": "every rank lights its own random ignition cells (bench_starts(rank)), so the ranks' fire "
                                  "fronts -- and with them the front-proportional step -- differ by a few per cent; `value` "
                                  "uses the slowest rank"},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pts_np.nbytes) * world,
                "d2h_bytes_per_step": (int(log_b * e2e_changes / e2e_steps) + 16 if not args.no_track else int(maps_np.nbytes)) * world,
                "steps": e2e_steps, "host_mirror_bytes": int(maps_np.nbytes) * world, "mirror_matches_download": e2e_ok,
                "host_mirror_memory": "pinned" if args.mirror == "pinned" else "pageable, MADV_HUGEPAGE",
                "ms_per_call": {"apply_points": e2e_calls[0], "step_enqueue": e2e_calls[1], "sync_fire_maps": e2e_calls[2]},
                "status_only": {"value": status_only_value, "unit": UNIT, "d2h_bytes_per_step": 16 * E * world,
                                "note": "the same loop when the observation stays in HBM (fire_map_device): points in, "
                                        "GameStatus / elapsed_time / update count of every env out"},
                "api": "FireEngine.apply_points (pinned H2D) + step + sync_fire_maps: every env's int8 fire_map is "
                       "brought up to date in host memory each step" + (" by patching the cells the device logged "
                       "as changed (%d B each)" % log_b if not args.no_track else " by a full download")},
        "gpu_launches": int(launches),
        "roofline": roof,
        "sanity": {"envs_running": running, "burned_cells_first_envs": burned, "steps_done_env0": int(nsteps[0])},
    }  # fmt: skip
    if age is not None:
        line["fire_age"] = age
    eng.close()  # the batch's device memory is not needed any more
    eng = None
    if world == 1 and not args.no_cpu_baseline:
        def gpu_maps(starts_, updates):  # envs lit at the same cells, stepped as a batch on the device
            with FireEngine(H, W, len(starts_), shared_static=True, device=local, **front_kw, **wl.engine_kwargs()) as one:
                one.set_static(wl.planes)
                one.reset(np.asarray(starts_, dtype=np.int32))
                one.step(updates)
                return one.fire_map()

        line["cpu_baseline"] = cpu_baseline_leg(args, wl, E, args.cpu_budget, gpu_maps if args.workload != "cfg3_perenv" else None)
    if world == 1 and unit_mode != "dense" and not args.no_dense_reference and args.workload != "cfg3_perenv":
        # the HBM-bound kernel of the design, measured live beside the front-proportional one: the same
        # batch stepped by the dense front end, i.e. the 1 B/cell TMA sweep over every cell
        try:  # an extra: it must never cost the run its line
            with make_engine(unit_skip=False, sweep_ldg=(args.sweep == "ldg")) as dense:
                dense.set_static(wl.planes)
                dense.reset(starts)
                dense.step(args.burn_in)
                dense.set_kernel_timing(True)
                dense.step(max(2, min(10, args.roofline_steps)))
                d_sweep, d_rows, d_eval, d_n = dense.kernel_ms()
                d_tasks, _ = dense.row_tasks()
            d_bytes = cells_rank * 1.0 + d_tasks * 8.0
            d_s = d_sweep / d_n * 1e-3
            line["roofline"]["dense_sweep"] = {
                "kernel": "k_sweep_" + args.sweep, "bound": "hbm", "ms_per_launch": d_s * 1e3, "bytes_per_launch": d_bytes,
                "achieved": d_bytes / d_s / 1e9, "peak": peak_gbs, "unit": "GB/s", "frac": d_bytes / d_s / 1e9 / peak_gbs,
                "traffic": load_traffic_note(args.workload, "k_sweep_" + args.sweep)[0],
                "traffic_source": "committed ncu capture (profiles/ncu_sweep_summary.json), same launch shape",
                "step_ms": (d_sweep + d_rows + d_eval) / d_n,
                "note": "the same batch stepped by the dense front end (every cell's state byte streamed once per step), at "
                        f"update {args.burn_in + 1}+: what slab mode uses; the crossover against the list-driven step is where "
                        "its k_front would take this long",
            }
        except Exception as exc:  # pragma: no cover
            line["roofline"]["dense_sweep"] = {"error": f"{type(exc).__name__}: {exc}"}
    print(json.dumps(line), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="target", choices=["target", "cfg3", "cfg3_perenv", "cfg2", "small", "cfg5", "cfg1"])
    ap.add_argument("--full-burn", action="store_true", help="cfg1 / cfg2: one env from ignition until GameStatus.QUIT")
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU")
    ap.add_argument("--burn-in", type=int, default=60)
    ap.add_argument("--rows-per-chunk", type=int, default=0)
    ap.add_argument("--env-groups", type=int, default=0, help="env groups stepped on separate streams (0 = auto)")
    ap.add_argument("--sweep", default="tma", choices=["tma", "ldg"], help="streaming front end of k_sweep")
    ap.add_argument("--front", default="auto", choices=["auto", "lists", "bits", "rows", "chunks", "dense"],
                    help="auto: the library's choice (row units for big batches, the dense sweep for small handles); "
                         "lists: the list-driven step; rows / chunks / dense: force a sweep front end")
    ap.add_argument("--age-curve", type=int, default=2000,
                    help="also step the batch from ignition to this many updates and report ms/step per block (0: skip)")
    ap.add_argument("--roofline-steps", type=int, default=20)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--parity-updates", type=int, default=150,
                    help="same-run parity: PARITY_ENVS are stepped this many updates on the device and on the host")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense-reference", action="store_true",
                    help="skip the extra pass that times the dense TMA sweep beside the default front end")
    ap.add_argument("--as-batch", action="store_true",
                    help="cfg5: step the 8192^2 grid as an ordinary one-env engine on one GPU (row units) instead of slab mode")
    ap.add_argument("--slab-sync", default="p2p", choices=["p2p", "nccl"], help="cfg5: how the slabs agree per step")
    ap.add_argument("--diag-back-to-back", action="store_true",
                    help="also time the next K steps right after the timed region (no barrier / idle GPU before them)")
    ap.add_argument("--no-track", action="store_true", help="e2e downloads every fire_map in full each step")
    ap.add_argument("--mirror", default="pinned", choices=["pinned", "thp"],
                    help="memory of the host fire_map mirror in the e2e phase")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.full_burn or args.workload == "cfg1":
        if args.workload not in ("cfg1", "cfg2"):
            raise SystemExit("--full-burn is defined for --workload cfg1 and cfg2")
        full_burn_arm(args)
    elif args.impl == "reference":
        reference_arm(args)
    elif args.workload == "cfg5" and not args.as_batch:
        gpu_arm_slab(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
