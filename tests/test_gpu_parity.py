"""
GPU parity tests: the CUDA stepper, driven through the C ABI (simfire_b200.FireEngine is a
one-to-one ctypes wrapper), against the golden trajectories recorded from the unmodified
reference and against the NumPy oracle on seeded inputs.

Tolerances.  fire_map / GameStatus / elapsed_time: bit-exact.  Rate of spread and the
float64 burn accumulator: 1e-5 relative (BASELINE.json north_star), plus an absolute term
for pairs where 1 + phi_w + phi_s cancels ("Mixed precision"): the
error of a float32 transcendental is relative to the un-cancelled magnitude, so the
absolute term is 3e-6 of the rate the same pair would have on flat ground.
"""
import os

import numpy as np
import pytest
from scenario_io import GOLDEN, check_trajectory, dense_params, load_scenario, scenario_names

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def engine_for(sc, **kw):
    from simfire_b200 import FireEngine

    eng = FireEngine(
        sc["H"], sc["W"], kw.pop("E", 1), pixel_scale=float(sc["ps"]), update_rate=float(sc["dt"]),
        max_fire_duration=int(sc["max_dur"]), max_time=sc["max_time"], attenuate_line_ros=bool(sc["attenuate"]),
        diagonal_spread=bool(sc["diagonal"]), M_f=float(sc["M_f"]), keep_ros=True, **kw,
    )  # fmt: skip
    eng.set_static(sc["planes"])
    return eng


class EngineAdapter:
    """Adapts FireEngine (env `env`) to scenario_io.check_trajectory."""

    def __init__(self, eng, env=0):
        self.eng, self.env = eng, env

    def apply_points(self, pts):
        if pts:
            self.eng.apply_points([(self.env, x, y, k) for x, y, k in pts])

    def step(self):
        self.eng.step(1)
        return int(self.eng.status()[0][self.env])

    def get_map(self):
        return self.eng.fire_map(self.env, 1)[0]

    def get_burn(self):
        return self.eng.plane("burn", self.env)

    def get_ros(self):
        return self.eng.plane("ros", self.env)

    def elapsed(self):
        return float(self.eng.status()[1][self.env])


def test_rate_of_spread_golden_pairs():
    """compute_rate_of_spread (rothermel.py:4-136) on the device vs the reference's output."""
    from simfire_b200 import rate_of_spread

    z = np.load(f"{GOLDEN}/rothermel_pairs.npz")
    h, S_T, S_e, p_p, M_f = z["consts"]
    args = [z[k] for k in ("w_0", "delta", "M_x", "sigma", "U", "U_dir")]
    R = rate_of_spread(z["direction"], *args, z["slope_mag"], z["slope_dir"], h=h, S_T=S_T, S_e=S_e, p_p=p_p, M_f=M_f)
    R_flat = rate_of_spread(z["direction"], *args, 0.0, 0.0, h=h, S_T=S_T, S_e=S_e, p_p=p_p, M_f=M_f)
    want = z["R"]
    assert np.array_equal(R == 0, want == 0)
    err = np.abs(R - want)
    tol = RTOL * np.abs(want) + 3e-6 * np.maximum(R_flat, np.abs(want))
    assert np.all(err <= tol), f"max excess {np.max(err - tol)} at {np.argmax(err - tol)}"
    # and within the north star's 1e-5 relative without the cancellation allowance for all but the two pairs
    # where 1 + phi_w + phi_s cancels: the count is pinned, not just bounded
    assert int(np.sum(err > RTOL * np.abs(want))) <= 2, int(np.sum(err > RTOL * np.abs(want)))


def test_rate_of_spread_known_answer():
    """simfire/world/_tests/test_rothermel.py:10-100 (src == dst there, so theta = 0 = direction 0)."""
    from simfire_b200 import rate_of_spread

    z = np.load(f"{GOLDEN}/rothermel_pairs.npz")
    R = rate_of_spread(np.zeros(8, np.int8), z["kat_w_0"], z["kat_delta"], z["kat_M_x"], z["kat_sigma"],
                       z["kat_U"], z["kat_U_dir"], 0.0, 0.0, M_f=0.03)  # fmt: skip
    np.testing.assert_array_almost_equal(R, z["kat_literal"], decimal=2)
    np.testing.assert_allclose(R, z["kat_R"], rtol=RTOL)


def _burn_tol(sc):
    # burn is a running sum of R*dt - attenuation: errors scale with the largest addend
    scale = max(float(np.max(np.abs(sc["burns"]))), float(np.max(np.abs(sc["ross"]))), 1.0)
    return dict(burn_exact=False, burn_rtol=RTOL, burn_atol=RTOL * scale)


@pytest.mark.parametrize("name", scenario_names())
def test_golden_trajectory(name):
    """fire_map bit-exact every step; burn / ros within tolerance; status and elapsed exact."""
    sc = load_scenario(name)
    with engine_for(sc) as eng:
        eng.reset([sc["init"]])
        check_trajectory(sc, EngineAdapter(eng), **_burn_tol(sc))


@pytest.mark.parametrize("name", ["scenario_a_models_diag_att", "scenario_d_midrun_mitigation",
                                  "scenario_b_models_4nbr_noatt", "scenario_f_max_time"])  # fmt: skip
@pytest.mark.parametrize("variant", ["wide_cells", "queue_overflow", "rows4", "rows64", "ldg", "ldg_wide", "ldg_rows8",
                                     "skip", "skip_wide", "skip_ldg_rows8", "skip_overflow", "noskip",
                                     "rowunits", "rowunits_wide", "rowunits_overflow",
                                     "lists", "lists_wide", "lists_overflow", "lists_no_rate_table",
                                     "bits", "bits_wide", "bits_overflow", "bits_groups3"])
def test_golden_trajectory_variants(name, variant, monkeypatch):
    """The 16-bit cell layout, the dense fallback taken on queue overflow, other chunk
    heights and the non-TMA streaming front end must give the same trajectories."""
    sc = load_scenario(name)
    kw = {"wide_cells": dict(wide_cells=True), "queue_overflow": dict(queue_capacity=3),
          "rows4": dict(rows_per_chunk=4), "rows64": dict(rows_per_chunk=64), "ldg": dict(sweep_ldg=True),
          "ldg_wide": dict(sweep_ldg=True, wide_cells=True),
          "ldg_rows8": dict(sweep_ldg=True, rows_per_chunk=8),
          # unit skipping (sweep only the flagged (env, rows, columns) units), forced on for these small grids
          "skip": dict(unit_skip=True, unit_chunks=True), "skip_wide": dict(unit_skip=True, unit_chunks=True, wide_cells=True),
          "skip_ldg_rows8": dict(unit_skip=True, unit_chunks=True, sweep_ldg=True, rows_per_chunk=8),
          "skip_overflow": dict(unit_skip=True, unit_chunks=True, queue_capacity=3), "noskip": dict(unit_skip=False),
          # ... with one row per unit: the flagged rows are the row tasks, nothing is swept
          "rowunits": dict(unit_skip=True), "rowunits_wide": dict(unit_skip=True, wide_cells=True),
          "rowunits_overflow": dict(unit_skip=True, queue_capacity=3),
          # the list-driven step (k_front / k_tail): one watch list instead of a sweep; a list that is too short
          # turns the handle to the dense form of the same per-cell routine
          "lists": dict(front_lists=True), "lists_wide": dict(front_lists=True, wide_cells=True),
          "lists_overflow": dict(front_lists=True, queue_capacity=3),
          "lists_no_rate_table": dict(front_lists=True),
          # the bitboard step (k_tiles + k_eval): bit planes per sprite duration, tiles of 32 rows x 30 columns
          "bits": dict(front_bits=True), "bits_wide": dict(front_bits=True, wide_cells=True),
          "bits_overflow": dict(front_bits=True, queue_capacity=3),
          "bits_groups3": dict(front_bits=True, env_groups=3)}[variant]  # fmt: skip
    if variant == "lists_no_rate_table":
        monkeypatch.setenv("SFB_NO_RTAB", "1")  # rates evaluated in the step instead of looked up
    with engine_for(sc, **kw) as eng:
        if variant.startswith("lists"):
            assert eng.unit_mode() == "lists"
        if variant.startswith("bits"):
            assert eng.unit_mode() == "bits"
        eng.reset([sc["init"]])
        check_trajectory(sc, EngineAdapter(eng), **_burn_tol(sc))
        if variant in ("queue_overflow", "lists_overflow", "bits_overflow"):
            assert eng.queue_stats()[1] == 3
        if variant == "lists_overflow":
            assert eng.queue_stats()[2]  # the handle went dense


def test_batched_envs_are_independent():
    """E copies of one terrain (shared static planes), different ignition points, each
    checked against its own NumPy-oracle run; env 2 is reset half way."""
    from oracle.dense_numpy import DenseFire

    sc = load_scenario("scenario_a_models_diag_att")
    H, W = sc["H"], sc["W"]
    rng = np.random.default_rng(7)
    burnable = np.argwhere(sc["planes"]["w_0"] > 0)
    starts = [tuple(int(v) for v in burnable[i][::-1]) for i in rng.choice(len(burnable), 5, replace=False)]
    E = len(starts)
    pre = [(e, x, y, k) for e in range(E) for x, y, k in sc["pre"] if (x, y) != starts[e]]
    with engine_for(sc, E=E, shared_static=True) as eng:
        eng.reset(starts)
        eng.apply_points(pre)
        oracles = []
        for e in range(E):
            o = DenseFire(sc["planes"], dense_params(sc), starts[e])
            o.apply_points([(x, y, k) for ee, x, y, k in pre if ee == e])
            oracles.append(o)
        for step in range(1, 41):
            if step == 15:  # RL-style reset of one env while the others keep running
                eng.reset([starts[0]], envs=[2])
                oracles[2] = DenseFire(sc["planes"], dense_params(sc), starts[0])
            eng.step(1)
            st, el, n = eng.status()
            maps = eng.fire_map()
            for e, o in enumerate(oracles):
                want_st = o.step()
                assert st[e] == want_st, (step, e)
                assert el[e] == o.elapsed_time, (step, e)
                assert n[e] == o.step_count, (step, e)
                assert np.array_equal(maps[e], o.status), f"env {e} step {step}"


@pytest.mark.parametrize("groups", [1, 2, 3, 5])
def test_env_groups_on_streams_equal_one_group(groups):
    """Stepping the envs as G groups on G streams is only a scheduling choice."""
    from simfire_b200 import FireEngine
    from simfire_b200.workloads import synthetic_operational

    wl = synthetic_operational(96, 160, seed=4, patch=8)
    E = 13
    kw = dict(wl.engine_kwargs(), attenuate_line_ros=True, max_time=60.0)
    starts = wl.burnable_starts(E, seed=3, margin=4)
    lines = [(e, x, 40 + e, 3 + (x % 3)) for e in range(E) for x in range(10, 150, 3)]
    res = []
    for g in (1, groups):
        with FireEngine(96, 160, E, shared_static=True, env_groups=g, track_changes=True, keep_ros=True, **kw) as eng:
            eng.set_static(wl.planes)
            eng.reset(starts)
            eng.apply_points(lines)
            mirror = np.zeros((E, 96, 160), dtype=np.int8)
            eng.sync_fire_maps(mirror)
            for _ in range(6):
                eng.step(7)
                eng.sync_fire_maps(mirror)
                assert np.array_equal(mirror, eng.fire_map())
            eng.reset(starts[:2], envs=[4, 9])
            eng.step(20, sync=False)
            eng.synchronize()
            res.append((eng.fire_map(), [eng.plane("burn", e) for e in (0, 6, 12)], eng.plane("ros", 12), eng.status()))
    a, b = res
    assert np.array_equal(a[0], b[0])
    assert all(np.array_equal(x, y) for x, y in zip(a[1], b[1])) and np.array_equal(a[2], b[2])
    assert all(np.array_equal(x, y) for x, y in zip(a[3], b[3]))


@pytest.mark.parametrize("front_end", ["tma", "ldg", "rows", "lists", "bits"])
@pytest.mark.parametrize("attenuate", [True, False])
def test_unit_skipping_changes_nothing(front_end, attenuate):
    """Looking only at the flagged units (chunks of rows that are then swept with either front end,
    or single rows that are the row tasks themselves) must give the same fire maps, burn planes,
    clocks and change logs as sweeping every unit, through ignitions that cross unit borders,
    control lines drawn mid-run, a map upload, a reset of some envs and envs that burn out; and
    it must really skip: far fewer units are listed than the handle has."""
    from oracle.dense_numpy import DenseFire, DenseParams
    from simfire_b200 import FireEngine
    from simfire_b200.workloads import synthetic_operational

    H, W, E = 150, 1100, 5  # three 512-cell strips, 19+ chunks of rows
    wl = synthetic_operational(H, W, seed=9, patch=16)
    kw = dict(wl.engine_kwargs(), attenuate_line_ros=attenuate, max_fire_duration=4)
    rng = np.random.default_rng(2)
    starts = wl.burnable_starts(E, seed=5, margin=3)
    starts[0] = (511, 75)   # on a strip border
    starts[1] = (1030, 6)
    lines0 = [(e, x, 60, 3 + (x % 3)) for e in range(E) for x in range(400, 700)]
    engines = []
    for skip in (False, True):
        if front_end == "lists" and skip:  # the list-driven step against the dense sweep
            eng = FireEngine(H, W, E, shared_static=True, front_lists=True, track_changes=True, **kw)
            assert eng.unit_mode() == "lists"
        elif front_end == "bits" and skip:  # the bitboard front end against the dense sweep
            eng = FireEngine(H, W, E, shared_static=True, front_bits=True, track_changes=True, env_groups=2, **kw)
            assert eng.unit_mode() == "bits"
        else:
            eng = FireEngine(H, W, E, shared_static=True, unit_skip=skip, unit_chunks=(front_end not in ("rows", "lists", "bits")),
                             sweep_ldg=(front_end == "ldg"), rows_per_chunk=8, track_changes=True, env_groups=2, **kw)  # fmt: skip
        eng.set_static(wl.planes)
        eng.reset(starts)
        eng.apply_points(lines0)
        engines.append(eng)
    oracle = DenseFire(wl.planes, DenseParams(**kw), tuple(int(v) for v in starts[0]))
    oracle.apply_points([(x, y, k) for e, x, y, k in lines0 if e == 0])
    mirrors = [np.zeros((E, H, W), np.int8) for _ in engines]
    listed = []
    try:
        for phase in range(6):
            n = 9
            for eng, mir in zip(engines, mirrors):
                eng.step(n)
                eng.sync_fire_maps(mir)
            for _ in range(n):
                oracle.step()
            a, b = engines[0].fire_map(), engines[1].fire_map()
            assert np.array_equal(a, b), f"phase {phase}: fire maps differ with unit skipping"
            assert np.array_equal(mirrors[1], b) and np.array_equal(mirrors[0], a)
            assert np.array_equal(a[0], oracle.status), f"phase {phase}: env 0 differs from the oracle"
            for e in (0, 3):
                assert np.array_equal(engines[0].plane("burn", e), engines[1].plane("burn", e))
            for x, y in zip(engines[0].status(), engines[1].status()):
                assert np.array_equal(x, y)
            listed.append(engines[1].unit_stats())
            assert engines[0].unit_stats()[0] == engines[0].unit_stats()[1]
            # between-step mutations of every kind
            if phase == 1:  # a control line near the fire of env 1 and one far from any fire
                pts = [(1, x, 12, 4) for x in range(900, 1090)] + [(2, x, 140, 5) for x in range(5, 60)]
                for eng in engines:
                    eng.apply_points(pts)
            if phase == 2:  # upload a map: burnt cells next to the front become fuel again
                m = engines[0].fire_map(3, 1)
                m[m == 2] = 0
                m[0, 100:110, 200:260] = 3
                for eng in engines:
                    eng.set_fire_map(m, env0=3)
            if phase == 3:
                for eng in engines:
                    eng.reset(np.array([[20, 20], [1000, 140]]), envs=[2, 4])
        total = listed[0][1]
        assert all(l < total // 4 for l, _ in listed), listed
    finally:
        for eng in engines:
            eng.close()


@pytest.mark.parametrize("fuse", [0, 1])
def test_bitboard_one_kernel_step_and_its_transitions(fuse, monkeypatch):
    """The bitboard step is k_tiles + k_eval.  With SFB_FUSE_EVAL=1 (opt-in: measured slower) a handle steps with
    ONE kernel (k_tiles closes the step itself) while no status has been written from outside since all envs
    were reset, and with two once control lines may exist.  Same results as the dense sweep either way and through
    the transitions: single and multi-step launches (graph replay), a control line (attenuation of untouched line
    cells included), a partial reset, a reset of every env, envs that burn out."""
    from simfire_b200 import FireEngine
    from simfire_b200.workloads import synthetic_operational

    monkeypatch.setenv("SFB_FUSE_EVAL", str(fuse))
    one = 1 if fuse else 2  # kernels per step while no control line can exist

    H, W, E = 70, 130, 4
    wl = synthetic_operational(H, W, seed=4, patch=8)
    kw = dict(wl.engine_kwargs(), attenuate_line_ros=True, max_fire_duration=3, pixel_scale=40.0)
    starts = wl.burnable_starts(E, seed=2, margin=2)
    with FireEngine(H, W, E, shared_static=True, front_bits=True, **kw) as a, \
            FireEngine(H, W, E, shared_static=True, unit_skip=False, **kw) as b:
        assert a.unit_mode() == "bits" and b.unit_mode() == "dense"
        for eng in (a, b):
            eng.set_static(wl.planes)
            eng.reset(starts)

        def same(what):
            assert np.array_equal(a.fire_map(), b.fire_map()), what
            for e in range(E):
                assert np.array_equal(a.plane("burn", e), b.plane("burn", e)), (what, e)
            for x, y in zip(a.status(), b.status()):
                assert np.array_equal(x, y), what

        def steps(n, chunks):
            l0 = a.launch_counts()[1]
            for k in chunks:
                a.step(k)
                b.step(k)
            assert sum(chunks) == n
            return (a.launch_counts()[1] - l0) / n

        assert steps(12, [1] * 5 + [7]) == one  # also through the two-step graph
        same("one-kernel steps")
        line = [(e, x, 20, 3 + e % 3) for e in range(E) for x in range(10, 120)]
        for eng in (a, b):
            eng.apply_points(line)
        assert steps(9, [1, 8]) == 2
        same("with a control line")
        for eng in (a, b):
            eng.reset(np.array([[5, 5], [100, 60]]), envs=[1, 3])
        assert steps(6, [6]) == 2  # envs 0 and 2 still carry their lines
        same("after a partial reset")
        for eng in (a, b):
            eng.reset(starts)
        assert steps(40, [3, 1, 36]) == one
        same("after a reset of every env")
        assert steps(200, [200]) == one  # until everything has burnt out
        same("burnt out")
        assert not a.status()[0].any()


@pytest.mark.parametrize("unit_skip", [False, True])
def test_grouped_multi_step_launch_equals_single_steps(unit_skip):
    """Multi-group handles can replay pairs of steps as one CUDA graph forked over the group streams
    (opt-in; change log off, n >= 4, from either parity); single steps are enqueued kernel by kernel.
    Same state either way, also across a control line drawn between two launches."""
    from simfire_b200 import FireEngine
    from simfire_b200.workloads import synthetic_operational

    wl = synthetic_operational(64, 200, seed=6, patch=8)
    E = 7
    kw = dict(wl.engine_kwargs(), attenuate_line_ros=True)
    starts = wl.burnable_starts(E, seed=1, margin=3)
    res = []
    for plan in ([1] * 27, [5, 9, 4, 1, 8]):
        with FireEngine(64, 200, E, shared_static=True, env_groups=3, unit_skip=unit_skip, keep_ros=True,
                        step_graph=True, **kw) as eng:
            eng.set_static(wl.planes)
            eng.reset(starts)
            done = 0
            for n in plan:
                eng.step(n)
                done += n
                if done == 14:
                    eng.apply_points([(e, x, 30, 3) for e in range(E) for x in range(20, 180)])
            assert done == 27
            res.append((eng.fire_map(), [eng.plane("burn", e) for e in (0, 3, 6)], eng.status()))
    a, b = res
    assert np.array_equal(a[0], b[0])
    assert all(np.array_equal(x, y) for x, y in zip(a[1], b[1]))
    assert all(np.array_equal(x, y) for x, y in zip(a[2], b[2]))


def _random_case(seed):
    """A seeded scenario: terrain, parameters, pre-drawn lines and a schedule of between-step
    actions (control lines, a map upload, a reset)."""
    from simfire_b200.workloads import synthetic_operational

    rng = np.random.default_rng(1000 + seed)
    H, W = int(rng.integers(9, 44)), int(rng.integers(9, 70))
    wl = synthetic_operational(H, W, seed=seed, patch=int(rng.integers(3, 9)), flat=bool(rng.integers(0, 2)))
    kw = dict(wl.engine_kwargs(), attenuate_line_ros=bool(rng.integers(0, 2)), diagonal_spread=bool(rng.integers(0, 4) > 0),
              max_fire_duration=int(rng.integers(1, 7)), pixel_scale=float(rng.choice([30.0, 50.0, 98.0, 150.0])),
              update_rate=float(rng.choice([0.5, 1.0, 2.5])))  # fmt: skip
    if rng.integers(0, 4) == 0:
        kw["max_time"] = float(rng.integers(5, 40))
    start = tuple(int(v) for v in wl.burnable_starts(1, seed=seed, margin=1)[0])
    n_lines = int(rng.integers(0, H * W // 8))
    lines = [(int(rng.integers(0, W)), int(rng.integers(0, H)), int(rng.integers(3, 6))) for _ in range(n_lines)]
    # FireSimulation.update_mitigation applies the fireline, scratchline and wetline managers in that
    # order (simulation.py:468-478): when two points of one call name a cell, the later KIND wins
    lines = sorted((pt for pt in lines if pt[:2] != start), key=lambda pt: pt[2])
    steps = int(rng.integers(20, 60))
    actions = {}
    for _ in range(int(rng.integers(0, 4))):
        t = int(rng.integers(1, steps))
        kind = int(rng.integers(0, 6))  # any BurnStatus may be written (mitigation.py:77)
        actions[t] = ("points", [(int(rng.integers(0, W)), int(rng.integers(0, H)), kind) for _ in range(int(rng.integers(1, 30)))])
    if rng.integers(0, 3) == 0:
        actions[int(rng.integers(1, steps))] = ("upload", int(rng.integers(0, 1 << 30)))
    return wl, kw, start, lines, steps, actions


@pytest.mark.parametrize("front", ["sweeps", "bits"])
@pytest.mark.parametrize("seed", range(24))
def test_random_scenarios_against_oracle(seed, front):
    """Seeded random terrains, parameters, control lines and between-step actions: fire_map,
    GameStatus and elapsed_time bit-exact against the NumPy oracle every step, the burn plane
    within tolerance.  The oracle runs first; a scenario in which some cell's accumulated burn
    comes within 2e-3 (relative to its increment) of the ignition threshold is not a fair test of
    bit-exactness ("Ties at the ignition threshold") and is skipped.  Odd
    seeds force row-unit skipping, seeds divisible by 4 chunk units, the rest sweep densely; "bits" runs
    every seed on the bitboard step (k_tiles + k_eval)."""
    from oracle.dense_numpy import DenseFire, DenseParams
    from simfire_b200 import FireEngine

    wl, kw, start, lines, steps, actions = _random_case(seed)
    ps = kw["pixel_scale"]
    oracle = DenseFire(wl.planes, DenseParams(**kw), start)
    oracle.apply_points(lines)
    rng = np.random.default_rng(seed)
    trace, uploads = [], {}
    for t in range(1, steps + 1):
        st = oracle.step()
        moved = oracle.ros != 0
        if st == 1 and moved.any():
            margin = np.abs(oracle.burn[moved] - ps) / np.maximum(np.abs(oracle.ros[moved]), 1.0)
            if margin.min() < 2e-3:
                pytest.skip(f"seed {seed}: burn within {margin.min():.1e} of the ignition threshold at step {t}")
        trace.append((st, oracle.status.copy(), oracle.burn.copy(), oracle.elapsed_time))
        if t in actions:
            what, arg = actions[t]
            if what == "points":
                oracle.apply_points(arg)
            else:
                m = oracle.status.copy()
                flip = np.random.default_rng(arg).random(m.shape) < 0.15
                m[flip & (m == 2)] = 0  # burnt cells become fuel again
                m[flip & (m == 0)] = 4
                uploads[t] = m
                oracle.set_fire_map(m)
    mode = dict(unit_skip=True) if seed % 2 else (dict(unit_skip=True, unit_chunks=True, rows_per_chunk=4) if seed % 4 == 0 else dict(unit_skip=False))
    if front == "bits":
        mode = dict(front_bits=True, wide_cells=bool(seed % 5 == 0))
    with FireEngine(wl.H, wl.W, 1, keep_ros=True, sweep_ldg=bool(seed % 3 == 0), **mode, **kw) as eng:
        assert front != "bits" or eng.unit_mode() == "bits"
        eng.set_static(wl.planes)
        eng.reset([start])
        if lines:
            eng.apply_points([(0, x, y, k) for x, y, k in lines])
        scale = max(1.0, max(float(np.max(np.abs(b))) for _, _, b, _ in trace))
        for t, (st, status, burn, elapsed) in enumerate(trace, start=1):
            eng.step(1)
            got_st, got_el, _ = eng.status()
            assert int(got_st[0]) == st, f"seed {seed} step {t}: GameStatus"
            assert np.array_equal(eng.fire_map(0, 1)[0], status), f"seed {seed} step {t}: fire_map"
            assert float(got_el[0]) == elapsed, f"seed {seed} step {t}: elapsed_time"
            if t % 7 == 0 or t == len(trace):
                np.testing.assert_allclose(eng.plane("burn", 0), burn, rtol=RTOL, atol=RTOL * scale)
            if t in actions:
                what, arg = actions[t]
                if what == "points":
                    eng.apply_points([(0, x, y, k) for x, y, k in arg])
                else:
                    eng.set_fire_map(uploads[t][None], env0=0)


def test_multi_step_launch_equals_single_steps():
    sc = load_scenario("scenario_c_random_fuel_hills")
    with engine_for(sc) as a, engine_for(sc) as b:
        for eng in (a, b):
            eng.reset([sc["init"]])
            eng.apply_points([(0, x, y, k) for x, y, k in sc["pre"]])
        for _ in range(12):
            a.step(1)
        b.step(12, sync=False)
        b.synchronize()
        assert np.array_equal(a.fire_map(), b.fire_map())
        assert np.array_equal(a.plane("burn"), b.plane("burn"))
        assert a.status()[1][0] == b.status()[1][0]


def test_update_with_host_maps():
    """sfb_update: the manager.update(fire_map) drop-in with host buffers."""
    sc = load_scenario("scenario_d_midrun_mitigation")
    with engine_for(sc) as eng:
        eng.reset([sc["init"]])
        fm = np.zeros((sc["H"], sc["W"]), dtype=np.int8)
        fm[sc["init"][1], sc["init"][0]] = 1
        for x, y, k in sc["pre"]:
            fm[y, x] = k
        map_at = {int(s): i for i, s in enumerate(sc["map_steps"])}
        for step in range(1, int(sc["n_steps"]) + 1):
            for x, y, k in sc["schedule"].get(step, ()):
                fm[y, x] = k  # the caller edits its own array, as mitigation.py:77 does
            st = eng.update(fm)
            assert st[0] == int(sc["status"][step - 1])
            if step in map_at:
                assert np.array_equal(fm, sc["maps"][map_at[step]]), f"step {step}"


@pytest.mark.parametrize("shape", [(64, 64), (50, 70)])  # pitch == W and pitch > W
def test_incremental_host_mirror_matches_full_download(shape):
    """sfb_sync_fire_maps: patching a host mirror from the device change log must give the
    same maps as downloading them, across steps, mitigation, resets and map replacement."""
    from simfire_b200 import FireEngine
    from simfire_b200.workloads import synthetic_operational

    H, W = shape
    wl = synthetic_operational(H, W, seed=5, patch=8)
    E = 6
    rng = np.random.default_rng(3)
    kw = dict(wl.engine_kwargs(), attenuate_line_ros=True)
    with FireEngine(H, W, E, shared_static=True, track_changes=True, **kw) as eng:
        eng.set_static(wl.planes)
        eng.reset(wl.burnable_starts(E, seed=2, margin=4))
        mirror = np.full((E, H, W), 77, dtype=np.int8)
        assert eng.sync_fire_maps(mirror) == -1  # first call: full download
        assert np.array_equal(mirror, eng.fire_map())
        patched = 0
        for it in range(40):
            if it % 3 == 0:
                pts = np.stack([rng.integers(0, E, 9), rng.integers(0, W, 9), rng.integers(0, H, 9),
                                rng.integers(3, 6, 9)], axis=1)  # fmt: skip
                eng.apply_points(pts)
            if it == 17:
                eng.reset(wl.burnable_starts(2, seed=9, margin=4), envs=[1, 4])
            eng.step(int(rng.integers(1, 4)))
            if it == 25:
                eng.set_fire_map(eng.fire_map(2, 1), env0=2)  # wholesale replacement -> full resync
                assert eng.sync_fire_maps(mirror) == -1
            else:
                n = eng.sync_fire_maps(mirror)
                assert n >= 0
                patched += n
            assert np.array_equal(mirror, eng.fire_map()), f"iteration {it}"
        assert patched > 100
        eng.set_tracking(False)  # paused: changes are not logged ...
        eng.step(3)
        eng.set_tracking(True)  # ... so the first sync after resuming downloads everything
        assert eng.sync_fire_maps(mirror) == -1
        assert np.array_equal(mirror, eng.fire_map())
        eng.step(2)
        assert eng.sync_fire_maps(mirror) >= 0
        assert np.array_equal(mirror, eng.fire_map())
        other = np.zeros_like(mirror)
        assert eng.sync_fire_maps(other) == -1  # a different buffer cannot be patched
        assert np.array_equal(other, mirror)


@pytest.mark.parametrize("groups", [1, 3])
def test_parallel_mirror_patching_paths(groups, monkeypatch):
    """The host side of sfb_sync_fire_maps on several threads, forced on for small logs: the
    one-pass path taken when a single step follows the between-step calls (its entries name every
    cell at most once) and the ordered two-pass path for everything else (several steps per sync, a
    mitigation call after the step, resets).  The patched mirror must always equal a download."""
    from simfire_b200 import FireEngine
    from simfire_b200.workloads import synthetic_operational

    monkeypatch.setenv("SFB_PATCH_PARALLEL_MIN", "8")
    monkeypatch.setenv("SFB_HOST_THREADS", "4")
    H, W, E = 70, 90, 9
    wl = synthetic_operational(H, W, seed=12, patch=8)
    kw = dict(wl.engine_kwargs(), attenuate_line_ros=True, max_fire_duration=3)
    rng = np.random.default_rng(4)
    with FireEngine(H, W, E, shared_static=True, env_groups=groups, track_changes=True, **kw) as eng:
        eng.set_static(wl.planes)
        eng.reset(wl.burnable_starts(E, seed=2, margin=3))
        mirror = np.zeros((E, H, W), np.int8)
        eng.sync_fire_maps(mirror)  # first call: full download
        patched = 0
        for it in range(40):
            pts = [(int(e), int(rng.integers(0, W)), int(rng.integers(0, H)), int(rng.integers(0, 6))) for e in range(E)]
            # a cell drawn twice in one call, and a cell next to the fire that may ignite in the same step
            pts += [(0, 5, 5, 3), (0, 5, 5, 5)]
            mode = it % 5
            if mode in (0, 1, 2):      # mitigation, one step, sync: the one-pass path
                eng.apply_points(pts)
                eng.step(1)
            elif mode == 3:            # several steps per sync: ordered path
                eng.apply_points(pts)
                eng.step(3)
            else:                      # a call after the step, and a reset of two envs: ordered path
                eng.step(1)
                eng.apply_points(pts)
                if it % 10 == 4:
                    eng.reset(wl.burnable_starts(2, seed=it, margin=3), envs=[1, 7])
            n = eng.sync_fire_maps(mirror)
            patched += max(0, n)
            assert np.array_equal(mirror, eng.fire_map()), f"iteration {it} (mode {mode})"
        assert patched > 2000  # the logs were really patched, not re-downloaded


@pytest.mark.parametrize("groups", [1, 2, 3, 4])
def test_mirror_sees_calls_made_after_the_step(groups):
    """Stream-ordering stress (round-1 hardware failure): step -> apply_points / reset -> sync.  The
    between-step kernels run on the handle's stream while the step ran on the group streams; the
    sync must read the log heads ordered after BOTH.  The handle's stream is stalled before the
    mitigation kernel so that a head read on the wrong stream would certainly miss its entries, the
    point batches are as large as the staged (no host sync) path allows, and the mirror must equal
    a download after every iteration."""
    from simfire_b200 import FireEngine
    from simfire_b200.workloads import synthetic_operational

    H, W, E = 96, 128, 12
    wl = synthetic_operational(H, W, seed=21, patch=8)
    kw = dict(wl.engine_kwargs(), attenuate_line_ros=True, max_fire_duration=3)
    rng = np.random.default_rng(8)
    iters = int(os.environ.get("SFB_STRESS_ITERS", "40" if os.environ.get("SFB_EMULATED") == "1" else "500"))
    with FireEngine(H, W, E, shared_static=True, env_groups=groups, track_changes=True, **kw) as eng:
        eng.set_static(wl.planes)
        eng.reset(wl.burnable_starts(E, seed=2, margin=3))
        mirror = np.zeros((E, H, W), np.int8)
        eng.sync_fire_maps(mirror)
        patched = 0
        for it in range(iters):
            n_pts = int(rng.integers(1, 4000))  # <= 64 KB of points: staged, returns without a sync
            pts = np.stack([rng.integers(0, E, n_pts), rng.integers(0, W, n_pts), rng.integers(0, H, n_pts),
                            rng.integers(0, 6, n_pts)], axis=1)  # fmt: skip
            eng.step(1, sync=False)
            if it % 3 == 0:
                eng.debug_stall(300)
            eng.apply_points(pts)
            if it % 50 == 49:  # everything burnt out or was painted over: start again in a few envs
                eng.reset(wl.burnable_starts(3, seed=it, margin=3), envs=[0, 5, 11])
            n = eng.sync_fire_maps(mirror)
            patched += max(0, n)
            assert np.array_equal(mirror, eng.fire_map()), f"iteration {it}"
        assert patched > iters * 100


def test_observation_tensor_matches_host_map():
    import torch

    sc = load_scenario("scenario_chaparral_64")
    with engine_for(sc) as eng:
        eng.reset([sc["init"]])
        eng.step(10)
        t = eng.fire_map_device()
        assert t.is_cuda and t.dtype == torch.int8 and tuple(t.shape) == (1, sc["H"], sc["W"])
        assert np.array_equal(t.cpu().numpy(), eng.fire_map())


def test_errors_are_reported():
    from simfire_b200 import FireEngine, SfbError

    with pytest.raises(SfbError):
        FireEngine(8, 8, 1, pixel_scale=50.0, update_rate=1.0, max_fire_duration=0)
    with FireEngine(8, 8, 1, pixel_scale=50.0, update_rate=1.0, max_fire_duration=4) as eng:
        with pytest.raises(SfbError):
            eng.reset([(8, 0)])
        with pytest.raises(SfbError):
            eng.apply_points([(0, 1, 1, 9)])
        with pytest.raises(SfbError):
            eng.plane("ros")
        with pytest.raises(ValueError):
            eng.set_static({k: np.zeros((3, 3)) for k in ("w_0", "delta", "M_x", "sigma", "U", "U_dir", "slope_mag", "slope_dir")})


@pytest.mark.parametrize("n_slabs,ldg_single", [(2, False), (3, True)])
def test_row_slabs_equal_single_grid(n_slabs, ldg_single):
    """Slab mode (cfg5) on one GPU: the grid split in row slabs, each slab an engine that reads
    its neighbours' edge rows through the halo pointers, must evolve exactly like one engine
    holding the whole grid -- including the env-wide flags (QUIT, early return, attenuation of
    untouched control lines) that are OR-ed across slabs every step."""
    from simfire_b200 import FireEngine
    from simfire_b200.slab import SlabGrid
    from simfire_b200.workloads import synthetic_operational

    H, W, E = 101, 160, 2
    wl = synthetic_operational(H, W, seed=8, patch=8)
    kw = dict(wl.engine_kwargs(), attenuate_line_ros=True, max_time=70.0)
    starts = [(80, 49), (30, 52)]  # next to the slab boundaries
    lines = [(e, x, y, 3 + (x % 3)) for e in range(E) for y in (20, 51, 75) for x in range(10, 150, 2)]
    grid = SlabGrid(H, W, wl.planes, n_slabs=n_slabs, E=E, **kw)
    with FireEngine(H, W, E, sweep_ldg=ldg_single, **kw) as ref:
        ref.set_static(wl.planes, env=-1)
        ref.reset(starts)
        ref.apply_points(lines)
        grid.reset(starts)
        grid.apply_points(lines)
        for it in range(30):
            n = 1 + it % 3
            ref.step(n)
            grid.step(n)
            assert np.array_equal(grid.fire_map(), ref.fire_map()), f"step block {it}"
            a, b = grid.status(), ref.status()
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), it
        assert (ref.fire_map() == 2).sum() > 500
        burn = np.concatenate([e.plane("burn", 1) for e in grid.engines], axis=0)
        assert np.array_equal(burn, ref.plane("burn", 1))
    grid.close()
