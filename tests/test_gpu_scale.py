"""
Parity at the BASELINE.json sizes.  Where the NumPy oracle finishes in seconds it is the
checker (cfg1 in full, one 512 x 512 env of cfg3); at the full batch sizes the checks are
size-independent properties: the independent implementations inside the library (TMA ring,
LDG register window, dense queue-overflow fallback, 16-bit cells) must agree bit for bit,
batched envs must equal single-env runs, translated ignitions on uniform terrain must give
translated fires, and the per-cell life cycle must be monotone.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle(wl, start):
    from oracle.dense_numpy import DenseFire, DenseParams

    return DenseFire(wl.planes, DenseParams(**wl.engine_kwargs()), tuple(int(v) for v in start))


def _run_against_oracle(wl, start, n_steps, check_every=1, **engine_kw):
    """Steps engine and oracle together; fire_map must be identical at every checked step.
    The oracle also yields the smallest relative distance of any ignition test from the
    threshold: float32 libm differences (<= 1e-6) cannot flip a test further away than that."""
    from simfire_b200 import FireEngine

    o = _oracle(wl, start)
    margin = np.inf
    with FireEngine(wl.H, wl.W, 1, **wl.engine_kwargs(), **engine_kw) as eng:
        eng.set_static(wl.planes)
        eng.reset([start])
        for step in range(1, n_steps + 1):
            before = o.burn.copy()
            st = o.step()
            changed = o.burn != before
            if changed.any():
                margin = min(margin, float(np.min(np.abs(o.burn[changed] - wl.pixel_scale)) / max(wl.pixel_scale, 1e-9)))
            eng.step(1)
            if margin < 2e-5:
                # an ignition test this close to the threshold could be flipped by a 1-ulp libm difference.
                # The inputs of these tests are seeded and were chosen (with the oracle, in the dev container)
                # so that this never happens: if it does, the input changed -- a failure, not a shorter pass.
                pytest.fail(f"test input: oracle ignition margin {margin:.1e} at step {step} of {n_steps} "
                            f"({step - 1} steps verified); choose another seed")
            if step % check_every == 0 or st != 1:
                gst, gel, gn = eng.status()
                assert int(gst[0]) == st and float(gel[0]) == o.elapsed_time and int(gn[0]) == o.step_count, step
                got = eng.fire_map(0, 1)[0]
                if not np.array_equal(got, o.status):
                    bad = np.argwhere(got != o.status)
                    raise AssertionError(f"fire_map differs at step {step} in {len(bad)} cells, first {bad[:4].tolist()}")
            if st != 1:
                break
        burn = eng.plane("burn")
    scale = max(1.0, float(np.max(np.abs(o.burn))))
    np.testing.assert_allclose(burn, o.burn, rtol=1e-5, atol=1e-5 * scale)
    return o


def test_cfg1_functional_flat_full_burn():
    """BASELINE configs[0]: 128 x 128 functional_config.yml, flat -- run to extinction."""
    from simfire_b200.workloads import cfg1_functional_flat

    wl = cfg1_functional_flat(128, start=(16, 16))
    o = _run_against_oracle(wl, wl.init_pos, 400)
    assert o.game_status == 0 and (o.status == 2).all()  # everything burned, then QUIT


def test_cfg3_single_env_512_against_oracle():
    """One env of BASELINE configs[2] (512 x 512 synthetic operational terrain with hills)."""
    from simfire_b200.workloads import synthetic_operational

    wl = synthetic_operational(512, 512, seed=0)
    o = _run_against_oracle(wl, wl.init_pos, 120, check_every=10)
    assert (o.status == 2).sum() > 2000


def test_cfg2_1024_against_oracle_short():
    """BASELINE configs[1]: 1024 x 1024 synthetic operational terrain, 1 env."""
    from simfire_b200.workloads import synthetic_operational

    wl = synthetic_operational(1024, 1024, seed=0)
    _run_against_oracle(wl, wl.init_pos, 40, check_every=20)


@pytest.mark.parametrize("front", ["lists", "rows", "bits"])
def test_target_batch_2048x1024_against_oracle(front):
    """The benchmarked configuration itself: 2048 x 2048 x 1024 envs on the bench terrain with the bench's
    ignition cells (bench.bench_starts), 150 updates, eight envs compared with the NumPy oracle -- each on
    the window its fire provably cannot leave in that many updates (oracle.reference_runner.window_around:
    cells outside it are UNBURNED in the full-grid run, which the test also checks on the device map).
    fire_map every 10 updates, status / elapsed_time / update count, and the float64 burn plane at the
    end.  The envs are bench.PARITY_ENVS: screened with tools/screen_parity_envs.py so that no ignition
    test comes within 2e-5 of the threshold (re-checked here: a smaller margin fails the test)."""
    from bench import PARITY_ENVS, bench_starts
    from oracle.dense_numpy import DenseFire, DenseParams
    from oracle.reference_runner import window_around
    from simfire_b200 import FireEngine
    from simfire_b200.workloads import synthetic_operational

    H = W = 2048
    E, n_updates = 1024, 150
    wl = synthetic_operational(H, W, seed=0, flat=True)
    starts = bench_starts(wl, E, 0)
    envs = PARITY_ENVS["target"]
    wins, oracles = [], []
    for e in envs:
        y0, x0, h, w = win = window_around(starts[e], n_updates, H, W)
        planes = {k: np.ascontiguousarray(np.broadcast_to(v, (H, W))[y0 : y0 + h, x0 : x0 + w]) for k, v in wl.planes.items()}
        oracles.append(DenseFire(planes, DenseParams(**wl.engine_kwargs()), (int(starts[e][0]) - x0, int(starts[e][1]) - y0)))
        wins.append(win)
    kw = {"lists": dict(front_lists=True), "rows": dict(unit_skip=True), "bits": dict(front_bits=True)}[front]
    margin = np.inf
    with FireEngine(H, W, E, shared_static=True, **kw, **wl.engine_kwargs()) as eng:
        assert eng.unit_mode() == front
        eng.set_static(wl.planes)
        eng.reset(starts)
        for block in range(n_updates // 10):
            eng.step(10)
            st, el, cnt = eng.status()
            for k, e in enumerate(envs):
                o = oracles[k]
                for _ in range(10):
                    before = o.burn.copy()
                    o.step()
                    ch = o.burn != before
                    if ch.any():
                        margin = min(margin, float(np.min(np.abs(o.burn[ch] - wl.pixel_scale)) / wl.pixel_scale))
                assert margin > 2e-5, f"test input: ignition margin {margin:.1e} in env {e}: re-screen bench.PARITY_ENVS"
                y0, x0, h, w = wins[k]
                got = eng.fire_map(e, 1)[0]
                assert np.array_equal(got[y0 : y0 + h, x0 : x0 + w], o.status), f"env {e}, update {10 * (block + 1)}"
                assert (got != 0).sum() == (o.status != 0).sum(), f"env {e}: cells outside the window changed"
                assert (int(st[e]), float(el[e]), int(cnt[e])) == (o.game_status, o.elapsed_time, o.step_count), e
        for k, e in enumerate(envs):
            y0, x0, h, w = wins[k]
            burn = eng.plane("burn", e)
            scale = max(1.0, float(np.max(np.abs(oracles[k].burn))))
            np.testing.assert_allclose(burn[y0 : y0 + h, x0 : x0 + w], oracles[k].burn, rtol=1e-5, atol=1e-5 * scale)
            assert np.count_nonzero(burn) == np.count_nonzero(oracles[k].burn)
        assert sum(int((o.status == 2).sum()) for o in oracles) > 50000  # the fires really grew


def _checksums(eng):
    """Per-env position-weighted checksum of the fire_map, computed on the device."""
    import torch

    t = eng.fire_map_device().to(torch.int64)
    E, H, W = t.shape
    w = (torch.arange(H * W, device=t.device, dtype=torch.int64) * 2654435761 % 1000003).view(1, H, W)
    return (t * w).sum(dim=(1, 2)).cpu().numpy(), (t == 2).sum(dim=(1, 2)).cpu().numpy(), (t == 1).sum(dim=(1, 2)).cpu().numpy()


@pytest.mark.parametrize("shape", [(512, 512, 1024), (2048, 2048, 128)])
def test_front_ends_agree_at_batch_size(shape):
    """cfg3 batch (512^2 x 1024 envs) and the target grid (2048^2): the list-driven step vs the TMA ring
    vs row units vs the LDG window vs 16-bit cells vs the dense form taken on a list overflow --
    identical fire maps for every env."""
    from simfire_b200 import FireEngine
    from simfire_b200.workloads import synthetic_operational

    H, W, E = shape
    wl = synthetic_operational(H, W, seed=0, flat=(H == 2048))
    kw = dict(wl.engine_kwargs(), attenuate_line_ros=True)
    starts = wl.burnable_starts(E, seed=77)
    rng = np.random.default_rng(1)
    lines = np.stack([rng.integers(0, E, 4 * E), rng.integers(0, W, 4 * E), rng.integers(0, H, 4 * E),
                      rng.integers(3, 6, 4 * E)], axis=1)  # fmt: skip
    # row units (the default at this size), the 16-bit layout, the dense TMA / LDG sweeps and the list-driven step
    variants = {"rows": dict(unit_skip=True), "default": {}, "wide": dict(wide_cells=True, unit_skip=True), "tma": dict(unit_skip=False),
                "lists": dict(front_lists=True),
                "lists_wide": dict(front_lists=True, wide_cells=True), "ldg": dict(sweep_ldg=True),
                "bits": dict(front_bits=True), "bits_wide_1group": dict(front_bits=True, wide_cells=True, env_groups=1)}
    if H == 512:
        variants["overflow"] = dict(queue_capacity=1000)
        variants["lists_overflow"] = dict(front_lists=True, queue_capacity=1000)
    results = {}
    for name, extra in variants.items():
        with FireEngine(H, W, E, shared_static=True, **kw, **extra) as eng:
            if name == "default":  # what a caller who asks for nothing gets at this size
                assert eng.unit_mode() == "bits"
            eng.set_static(wl.planes)
            eng.reset(starts)
            eng.apply_points(lines)
            seq = []
            for _ in range(3):
                eng.step(20 if "overflow" not in name else 6)
                seq.append(_checksums(eng))
            results[name] = (seq, eng.status())
            if "overflow" in name:
                assert eng.queue_stats()[2]  # the last step really overflowed / the list handle went dense
    if "overflow" in results:  # the fallbacks are slow: compare them over their shorter run
        with FireEngine(H, W, E, shared_static=True, **kw) as eng:
            eng.set_static(wl.planes)
            eng.reset(starts)
            eng.apply_points(lines)
            seq = []
            for _ in range(3):
                eng.step(6)
                seq.append(_checksums(eng))
            for nm in ("overflow", "lists_overflow"):
                for a, b in zip(seq, results.pop(nm)[0]):
                    assert all(np.array_equal(x, y) for x, y in zip(a, b)), nm
    base_seq, base_st = results["rows"]
    assert base_seq[-1][1].sum() > 50 * E  # fires really spread
    for name, (seq, st) in results.items():
        for k, (a, b) in enumerate(zip(seq, base_seq)):
            assert all(np.array_equal(x, y) for x, y in zip(a, b)), f"{name} differs from row units after block {k}"
        assert all(np.array_equal(x, y) for x, y in zip(st, base_st)), name


def test_batched_envs_equal_single_env_runs_at_cfg3_size():
    from simfire_b200 import FireEngine
    from simfire_b200.workloads import synthetic_operational

    H = W = 512
    E = 1024
    wl = synthetic_operational(H, W, seed=0)
    starts = wl.burnable_starts(E, seed=5)
    with FireEngine(H, W, E, shared_static=True, **wl.engine_kwargs()) as eng:
        eng.set_static(wl.planes)
        eng.reset(starts)
        eng.step(60)
        st, el, n = eng.status()
        for e in (0, 511, 1023):
            with FireEngine(H, W, 1, **wl.engine_kwargs()) as one:
                one.set_static(wl.planes)
                one.reset([starts[e]])
                one.step(60)
                assert np.array_equal(one.fire_map(0, 1)[0], eng.fire_map(e, 1)[0]), e
                assert np.array_equal(one.plane("burn"), eng.plane("burn", e)), e
                s1, e1, n1 = one.status()
                assert (s1[0], e1[0], n1[0]) == (st[e], el[e], n[e])


def test_translation_invariance_on_uniform_terrain_2048():
    """Uniform fuel, wind and flat ground on the 2048^2 target grid: moving the ignition point
    moves the whole fire (maps and burn accumulators) by the same offset."""
    from simfire_b200 import FireEngine
    from simfire_b200.workloads import cfg1_functional_flat

    wl = cfg1_functional_flat(2048)
    starts = [(700, 900), (1213, 517), (1500, 1501)]  # odd offsets: different strips, chunks, lanes
    kw = dict(wl.engine_kwargs(), pixel_scale=300.0, max_time=None)  # several steps per cell in the slow directions
    with FireEngine(wl.H, wl.W, len(starts), shared_static=True, **kw) as eng:
        eng.set_static(wl.planes)
        eng.reset(starts)
        eng.step(150)
        maps = eng.fire_map()
        r = 200
        ref_map = ref_burn = None
        for e, (x, y) in enumerate(starts):
            win = maps[e, y - r : y + r, x - r : x + r]
            burn = eng.plane("burn", e)[y - r : y + r, x - r : x + r]
            assert (maps[e] != 0).sum() == (win != 0).sum()  # the fire is inside the window
            if e == 0:
                ref_map, ref_burn = win, burn
                assert (win == 2).sum() > 3000
            else:
                assert np.array_equal(win, ref_map), e
                assert np.array_equal(burn, ref_burn), e


def test_cell_life_cycle_is_monotone():
    """UNBURNED -> BURNING -> BURNED only (no mitigation): burned counts never decrease, a
    cell never goes back, BURNING cells are exactly the cells carrying a sprite."""
    from simfire_b200 import FireEngine
    from simfire_b200.workloads import synthetic_operational

    wl = synthetic_operational(512, 512, seed=2)
    E = 16
    with FireEngine(512, 512, E, shared_static=True, **wl.engine_kwargs()) as eng:
        eng.set_static(wl.planes)
        eng.reset(wl.burnable_starts(E, seed=3))
        prev = eng.fire_map()
        for _ in range(12):
            eng.step(10)
            cur = eng.fire_map()
            assert cur.max() <= 2 and cur.min() >= 0
            assert np.all(cur >= prev)  # 0 -> 1 -> 2 only
            age = eng.plane("age", 3)
            assert np.array_equal(age >= 0, cur[3] == 1)
            assert age.max() <= wl.max_fire_duration  # == max: pruned by the next update()
            prev = cur


@pytest.mark.parametrize("shape,start", [((1, 1), (0, 0)), ((1, 40), (39, 0)), ((40, 1), (0, 0)), ((3, 3), (2, 2)),
                                          ((17, 513), (512, 16)), ((70, 1030), (0, 69))])
@pytest.mark.parametrize("front", ["default", "bits"])
def test_degenerate_and_ragged_grids_against_oracle(shape, start, front):
    """Single cells, single rows / columns, widths that are not a multiple of the 16-cell load or
    of the 512-cell strip, ignition in a corner: the reference's bounds handling (fire.py:192-205)."""
    from simfire_b200.workloads import Workload

    H, W = shape
    rng = np.random.default_rng(H * 1000 + W)
    planes = dict(w_0=rng.uniform(0.02, 0.4, (H, W)), delta=rng.uniform(0.5, 4.0, (H, W)),
                  M_x=rng.uniform(0.15, 0.4, (H, W)), sigma=rng.uniform(1200, 3400, (H, W)),
                  U=rng.uniform(0, 2000, (H, W)), U_dir=rng.uniform(0, 360, (H, W)),
                  slope_mag=rng.uniform(0, 0.3, (H, W)), slope_dir=rng.uniform(-3, 3, (H, W)))  # fmt: skip
    planes["w_0"][rng.random((H, W)) < 0.1] = 0.0
    planes["w_0"][start[1], start[0]] = 0.2
    wl = Workload("ragged", H, W, planes, pixel_scale=60.0, update_rate=1.5, max_fire_duration=3, max_time=90.0,
                  attenuate_line_ros=True, diagonal_spread=True, M_f=0.02, init_pos=start)  # fmt: skip
    _run_against_oracle(wl, start, 70, **(dict(front_bits=True) if front == "bits" else {}))


@pytest.mark.parametrize("max_dur", [1, 30, 31, 200])
def test_fire_duration_limits_of_the_two_cell_layouts(max_dur):
    """max_fire_duration 30 is the last value the 8-bit cell can encode, 31 switches to 16-bit."""
    from simfire_b200.workloads import synthetic_operational

    wl = synthetic_operational(64, 96, seed=13, patch=8)  # (seed 13: ignition margin > 3e-4 over all 260 updates)
    wl.max_fire_duration = max_dur
    wl.pixel_scale = 400.0  # slow spread: sprites really live for many steps
    _run_against_oracle(wl, wl.init_pos, 90 if max_dur < 100 else 260, check_every=5)


@pytest.mark.parametrize("max_dur", [1, 2, 7])
def test_fire_duration_limits_of_the_bitboard_ring(max_dur):
    """One sprite plane per duration: a ring of 2 (max_fire_duration 1) up to the 8 planes the tile kernel
    holds in registers; longer-lived sprites fall back to the byte front ends."""
    from simfire_b200 import FireEngine
    from simfire_b200.workloads import synthetic_operational

    wl = synthetic_operational(64, 96, seed=13, patch=8)
    wl.max_fire_duration = max_dur
    wl.pixel_scale = 400.0
    _run_against_oracle(wl, wl.init_pos, 90, check_every=5, front_bits=True)
    with FireEngine(64, 96, 1, front_bits=True, **dict(wl.engine_kwargs(), max_fire_duration=8)) as eng:
        assert eng.unit_mode() != "bits"
