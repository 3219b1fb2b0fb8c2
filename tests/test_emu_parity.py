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
NEEDS_TORCH_CUDA = "not observation_tensor and not row_slabs and not device_view"  # (device_view: torch tensors over raw pointers)


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


def test_experimental_k_rows_variant_under_emulation():
    """-DSFB_ROWS_V2 (one REDUX for the group mask, 16-byte pad vectors instead of single halo cells,
    per-env base pointers cached): not the default -- it needs 12-16 more registers and has not been
    measured on a B200 yet -- but it has to stay correct so that it can be A/B'd."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    from build_emu import build_variant

    env = dict(os.environ, SFB_LIB=build_variant("SFB_ROWS_V2"), SFB_EMULATED="1")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "tests/test_gpu_parity.py",
           "-k", "golden_trajectory or random_scenarios or unit_skipping"]  # fmt: skip
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, "\n".join((res.stdout + res.stderr).splitlines()[-25:])


def test_parity_under_address_sanitizer():
    """The emulator build with -fsanitize=address and exact-size "device" allocations: any kernel or
    host access past the end of a plane, work list, row-task list, flag array or change log aborts.
    (A subset here to keep the CPU suite short; the whole emulated GPU suite, also with unit
    skipping forced on for every handle, is clean: profiles/r01b_emu_asan.txt.)"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    from build_emu import asan_runtime, build_asan

    rt = asan_runtime()
    if not os.path.exists(rt):
        pytest.skip("libasan is not installed")
    env = dict(os.environ, SFB_LIB=build_asan(), SFB_EMULATED="1", LD_PRELOAD=rt,
               ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1")  # fmt: skip
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "tests/test_gpu_parity.py",
           "-k", "golden_trajectory or parallel_mirror or grouped_multi_step"]  # fmt: skip
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, "\n".join((res.stdout + res.stderr).splitlines()[-25:])


def _emu_env():
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    from build_emu import build

    return dict(os.environ, SFB_LIB=build(), SFB_EMULATED="1")


def test_bench_gpu_arm_dry_run_under_emulation():
    """bench.py's GPU arm, every phase (device-timed steps, per-kernel pass, e2e with the patched
    host mirror), on a tiny workload against the emulator: catches Python-level mistakes in the
    bench before the GPU box does.  The numbers mean nothing; the line's shape is checked."""
    import json

    for front in ("lists", "bits", "rows", "dense"):
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "bench_dry_run.py"), "--workload", "small",
                              "--steps", "3", "--warmup", "3", "--burn-in", "5", "--roofline-steps", "2", "--e2e-steps", "3",
                              "--age-curve", "50" if front == "lists" else "0", "--parity-updates", "12",
                              *(["--no-cpu-baseline"] if front == "dense" else ["--cpu-budget", "1"]), "--front", front],
                             cwd=ROOT, env=_emu_env(), capture_output=True, text=True, timeout=600)  # fmt: skip
        assert res.returncode == 0, res.stderr[-2000:]
        line = json.loads(res.stdout.strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert key in line, key
        assert line["e2e"]["mirror_matches_download"] is True
        assert line["gpu_launches"] > 0
        rf = line["roofline"]
        for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel_ms_per_launch", "front"):
            assert key in rf, key
        assert rf["front"] == front
        if front == "lists":
            assert rf["kernel"] == "k_front" and rf["per_launch"]["examined"] > 0 and not rf["went_dense"]
            assert line["gpu_launches"] == 3  # one kernel per step (the small workload has no attenuation)
            assert line["fire_age"]["updates"] == 50 and len(line["fire_age"]["blocks"]) == 1
        if front == "bits":
            assert rf["kernel"] in ("k_tiles", "k_eval") and rf["per_launch"]["candidates"] > 0
        if front == "dense":
            assert rf["unit_skipping"]["units_listed"] == rf["unit_skipping"]["units_total"]
        else:  # the CPU leg also checks the device against the reference (or the port) on the envs it names
            cb = line["cpu_baseline"]
            assert cb["kind"] in ("reference", "port") and cb["parity"]["full_grid"]["fire_map_equal"] is True
            assert cb["parity"]["windows"]["all_equal"] is True and cb["parity"]["windows"]["updates"] == 12
            ds = rf["dense_sweep"]  # ... and the dense TMA sweep is timed beside the default front end
            assert ds["bound"] == "hbm" and ds["kernel"].startswith("k_sweep") and ds["bytes_per_launch"] >= 256 * 256 * 64


def test_bench_full_burn_dry_run_under_emulation():
    """`bench.py --workload cfg1 --full-burn` (BASELINE config 1 until GameStatus.QUIT), checked
    against the oracle inside the bench itself."""
    import json

    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "bench_dry_run.py"), "--workload", "cfg1",
                          "--full-burn"], cwd=ROOT, env=_emu_env(), capture_output=True, text=True, timeout=600)  # fmt: skip
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["result"]["unburned_cells"] == 0 and line["result"]["burned_cells"] == 128 * 128
    assert line["cpu_baseline"]["parity"] == {"fire_map_equal": True, "updates_equal": True}
    assert line["cpu_baseline"]["kind"] in ("reference", "port")


def test_smoke_under_emulation():
    res = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, env=_emu_env(),
                         capture_output=True, text=True, timeout=600)  # fmt: skip
    assert res.returncode == 0 and "smoke ok" in res.stdout, res.stdout[-1000:] + res.stderr[-2000:]


def test_log_patching_orders_entries_of_one_cell():
    """The host side of sfb_sync_fire_maps (apply_log in sfb.cu, called directly through the emulator
    build's test hook) on adversarial logs: many entries per cell with alternating states and env
    resets in between must end as if applied one by one in log order (the two-pass path on several
    threads); a setup prefix followed by unique cells must as well (the one-pass path)."""
    import ctypes as C

    import numpy as np

    lib = C.CDLL(_emu_env()["SFB_LIB"])
    lib.sfb_emu_bench_apply_log.restype = C.c_double
    H, W, E = 37, 48, 6  # pitch == W (linear plane)
    rng = np.random.default_rng(3)
    n = 60000
    cells = rng.integers(0, E * H * W, n).astype(np.uint64)
    hot = rng.integers(0, E * H * W, 40).astype(np.uint64)  # a few cells that change again and again
    pick = rng.random(n) < 0.4
    cells[pick] = hot[rng.integers(0, len(hot), int(pick.sum()))]
    states = rng.integers(0, 6, n).astype(np.uint64)
    log = cells | (states << np.uint64(48))
    resets = np.sort(rng.choice(n, 25, replace=False))
    log[resets] = rng.integers(0, E, len(resets)).astype(np.uint64) | (np.uint64(7) << np.uint64(48))  # LOG_ENV_RESET
    want = np.full(E * H * W, 9, np.int8)
    for e in log:
        idx, st = int(e) & 0xFFFFFFFFFFFF, (int(e) >> 48) & 7
        if st == 7:
            want[idx * H * W : (idx + 1) * H * W] = 0
        else:
            want[idx] = st
    for threads in (1, 3, 4):
        got = np.full(E * H * W, 9, np.int8)
        lib.sfb_emu_bench_apply_log(H, W, E, C.c_void_p(log.ctypes.data), C.c_longlong(n), C.c_void_p(got.ctypes.data), threads, 1, 0)
        assert np.array_equal(got, want), f"ordered path, {threads} threads"
    # one-pass path: setup entries (bit 56) first, duplicates allowed among them; then every cell at most once
    uniq = rng.permutation(E * H * W)[:30000].astype(np.uint64)
    step = uniq | (rng.integers(1, 3, len(uniq)).astype(np.uint64) << np.uint64(48))
    setup_cells = np.concatenate([uniq[:200], uniq[:200]])  # also cells the step changes afterwards
    setup = setup_cells | (rng.integers(0, 6, len(setup_cells)).astype(np.uint64) << np.uint64(48)) | (np.uint64(1) << np.uint64(56))
    log2 = np.concatenate([setup, step])
    want2 = np.full(E * H * W, 9, np.int8)
    for e in log2:
        want2[int(e) & 0xFFFFFFFFFFFF] = (int(e) >> 48) & 7
    for threads in (2, 4):
        got = np.full(E * H * W, 9, np.int8)
        lib.sfb_emu_bench_apply_log(H, W, E, C.c_void_p(log2.ctypes.data), C.c_longlong(len(log2)), C.c_void_p(got.ctypes.data), threads, 1, 1)
        assert np.array_equal(got, want2), f"one-pass path, {threads} threads"
    # padded rows (pitch 64 > W 50): log indices address the padded plane, the mirror is dense
    H, W, E, pitch = 21, 50, 5, 64
    n = 40000
    env, y, x = rng.integers(0, E, n), rng.integers(0, H, n), rng.integers(0, W, n)
    st = rng.integers(0, 6, n)
    log3 = ((env * H + y) * pitch + x).astype(np.uint64) | (st.astype(np.uint64) << np.uint64(48))
    want3 = np.full((E, H, W), 9, np.int8)
    for e_, y_, x_, s_ in zip(env, y, x, st):
        want3[e_, y_, x_] = s_
    for threads in (1, 4):
        got = np.full((E, H, W), 9, np.int8)
        lib.sfb_emu_bench_apply_log(H, W, E, C.c_void_p(log3.ctypes.data), C.c_longlong(n), C.c_void_p(got.ctypes.data), threads, 1, 0)
        assert np.array_equal(got, want3), f"padded rows, {threads} threads"
