"""
GPU tests of the reference-shaped Python surface: the `RothermelFireManager` drop-in
(simfire/game/managers/fire.py) and `FireSimulation` (simfire/sim/simulation.py), replayed
against vectors recorded from the unmodified reference (tests/golden/gen_golden.py,
tests/golden/gen_api_golden.py) and against re-statements of the reference's own unit tests
(simfire/game/managers/_tests/test_fire.py, simfire/sim/_tests/test_simulation.py).
"""
import ast
import types

import numpy as np
import pytest
import yaml
from scenario_io import GOLDEN, load_scenario

pytestmark = pytest.mark.gpu


def _terrain(H, W, fuels, elevations):
    return types.SimpleNamespace(fuels=fuels, elevations=elevations, screen_size=(H, W))


def _manager_for(sc, **kw):
    from simfire_b200.fire_manager import RothermelFireManager
    from simfire_b200.parameters import Environment, FuelParticle

    p = sc["planes"]
    fuels = np.stack([p["w_0"], p["delta"], p["M_x"], p["sigma"]], axis=-1)
    return RothermelFireManager(
        sc["init"], 2, int(sc["max_dur"]), float(sc["ps"]), float(sc["dt"]), FuelParticle(),
        _terrain(sc["H"], sc["W"], fuels, sc["elevations"]), Environment(float(sc["M_f"]), p["U"], p["U_dir"]),
        max_time=sc["max_time"], attenuate_line_ros=bool(sc["attenuate"]), headless=True,
        diagonal_spread=bool(sc["diagonal"]), **kw,
    )  # fmt: skip


@pytest.mark.parametrize("name", ["scenario_c_random_fuel_hills", "scenario_d_midrun_mitigation",
                                  "scenario_f_max_time"])  # fmt: skip
def test_manager_dropin_replays_reference(name):
    """`fire_map, status = manager.update(fire_map)` with the caller's int64 map, edited in
    place between calls exactly as `ControlLineManager.update` does (mitigation.py:77)."""
    from simfire_b200.enums import BurnStatus, GameStatus

    sc = load_scenario(name)
    mgr = _manager_for(sc, keep_rate_of_spread=True)
    # slopes are computed by the manager from the elevations, as fire.py:436-449 does
    assert np.array_equal(mgr.slope_mag, sc["planes"]["slope_mag"])
    assert np.array_equal(mgr.slope_dir, sc["planes"]["slope_dir"])
    fire_map = np.full((sc["H"], sc["W"]), BurnStatus.UNBURNED)
    assert fire_map.dtype == np.int64
    fire_map[sc["init"][1], sc["init"][0]] = BurnStatus.BURNING
    for x, y, k in sc["pre"]:
        fire_map[y, x] = k
    map_at = {int(s): i for i, s in enumerate(sc["map_steps"])}
    for step in range(1, int(sc["n_steps"]) + 1):
        for x, y, k in sc["schedule"].get(step, ()):
            fire_map[y, x] = k
        out, status = mgr.update(fire_map)
        assert out is fire_map and isinstance(status, GameStatus)
        assert int(status) == int(sc["status"][step - 1]), step
        assert mgr.elapsed_time == float(sc["elapsed"][step - 1]), step
        if step in map_at:
            assert np.array_equal(fire_map, sc["maps"][map_at[step]]), f"{name} step {step}"
    np.testing.assert_allclose(mgr.burn_amounts, sc["burns"][-1], rtol=1e-5,
                               atol=1e-5 * float(np.max(np.abs(sc["burns"]))))  # fmt: skip
    mgr.close()


def _simple_manager(size=11, pixel_scale=1e-3, max_fire_duration=4, **kw):
    from simfire_b200.config import chaparral
    from simfire_b200.fire_manager import RothermelFireManager
    from simfire_b200.parameters import Environment, FuelParticle

    fuels = np.empty((size, size), dtype=object)
    fuels.fill(chaparral(1113))
    terrain = _terrain(size, size, fuels, np.zeros((size, size)))
    env = Environment(0.03, kw.pop("U", 88.0), kw.pop("U_dir", 135.0))
    return RothermelFireManager((size // 2, size // 2), 2, max_fire_duration, pixel_scale, 1.0, FuelParticle(),
                                terrain, env, headless=True, **kw)  # fmt: skip


def test_update_ignites_all_neighbours_when_pixel_scale_is_tiny():
    """test_fire.py:326-392 (which zeroes pixel_scale after construction): after one update
    every in-bounds neighbour burns (8 or 4)."""
    for diagonal, want in ((True, 9), (False, 5)):
        mgr = _simple_manager(diagonal_spread=diagonal)
        fm = np.zeros((11, 11), dtype=np.int64)
        fm[5, 5] = 1
        fm, status = mgr.update(fm)
        assert int(status) == 1
        assert (fm == 1).sum() == want
        assert sorted(mgr.sprites) == sorted((int(x), int(y)) for y, x in np.argwhere(fm == 1))
        assert mgr.elapsed_time == 1.0
        mgr.close()


def test_new_locations_bounds_and_blocked_neighbours():
    """test_fire.py:61-122 (`_get_new_locs`): a fire in a corner only reaches its in-bounds
    neighbours, and a neighbour that is BURNED (or BURNING) is not a destination while UNBURNED
    and the three control-line kinds are (`_filter_function`, fire.py:192-205).  The reference
    test calls `_get_new_locs` directly; here one update with a tiny pixel_scale ignites exactly
    the destinations, read back from the new Fire sprites in the reference's order
    (sorted by (y, x), fire.py:566)."""
    from simfire_b200.config import chaparral
    from simfire_b200.fire_manager import RothermelFireManager
    from simfire_b200.parameters import Environment, FuelParticle

    size = 9

    def run(init, blocked=(), lines=(), diagonal=True):
        fuels = np.empty((size, size), dtype=object)
        fuels.fill(chaparral(1113))
        mgr = RothermelFireManager(init, 2, 4, 1e-3, 1.0, FuelParticle(), _terrain(size, size, fuels, np.zeros((size, size))),
                                   Environment(0.03, 88.0, 135.0), headless=True, diagonal_spread=diagonal)  # fmt: skip
        fm = np.zeros((size, size), dtype=np.int64)
        fm[init[1], init[0]] = 1
        for (x, y), k in list(blocked) + list(lines):
            fm[y, x] = k
        fm, _ = mgr.update(fm)
        new = [xy for xy in mgr.sprites if tuple(xy) != tuple(init)]
        mgr.close()
        return new, fm

    # too small: (0, 0) -> (x+1, y), (x+1, y+1), (x, y+1); listed in (y, x) order
    new, _ = run((0, 0))
    assert new == [(1, 0), (0, 1), (1, 1)]
    # far corner -> its three in-bounds neighbours
    new, _ = run((size - 1, size - 1))
    assert new == [(size - 2, size - 2), (size - 1, size - 2), (size - 2, size - 1)]
    # the cell at (x+1, y) is BURNED: all 8-connected points except that one
    x = y = size // 2
    new, fm = run((x, y), blocked=[((x + 1, y), 2)])
    want = {(x + 1, y + 1), (x, y + 1), (x - 1, y + 1), (x - 1, y), (x - 1, y - 1), (x, y - 1), (x + 1, y - 1)}
    assert set(new) == want and len(new) == 7 and fm[y, x + 1] == 2
    # control lines of every kind are destinations (and ignite once their burn exceeds pixel_scale)
    new, fm = run((x, y), lines=[((x + 1, y), 3), ((x, y + 1), 4), ((x - 1, y), 5)], diagonal=False)
    assert set(new) == {(x, y - 1)}  # the three lines were attenuated below the threshold ...
    assert fm[y, x + 1] == 3 and fm[y + 1, x] == 4 and fm[y, x - 1] == 5  # ... and stay lines
    # 4-neighbour variant (fire.py:223-228)
    new, _ = run((x, y), diagonal=False)
    assert set(new) == {(x + 1, y), (x, y + 1), (x - 1, y), (x, y - 1)}


def test_wind_conversion():
    """test_fire.py:205-260: wind speed / direction given as a float, a nested sequence or an
    array all become (H, W) arrays on the manager."""
    size = 11
    for U, U_dir in ((7.0, 90.0), ([[7.0] * size for _ in range(size)], [[90.0] * size for _ in range(size)]),
                     (np.full((size, size), 7.0), np.full((size, size), 90.0))):  # fmt: skip
        mgr = _simple_manager(size=size, U=U, U_dir=U_dir)
        assert isinstance(mgr.U, np.ndarray) and isinstance(mgr.U_dir, np.ndarray)
        assert mgr.U.shape == (size, size) and mgr.U_dir.shape == (size, size)
        assert np.all(mgr.U == 7.0) and np.all(mgr.U_dir == 90.0)
        mgr.close()


def test_mitigation_points_land_in_fire_map_and_run_semantics():
    """test_mitigation.py:65-100 (the points given to a control-line manager are exactly the cells
    of that kind in the fire_map) and test_simulation.py:84-121 (`run("1h")` leaves burnt cells,
    after `reset()` one update advances `elapsed_time` by `update_rate`)."""
    from simfire_b200.config import Config
    from simfire_b200.simulation import FireSimulation

    z = np.load(f"{GOLDEN}/api_sequence_flat64.npz")
    cfg = yaml.safe_load(str(z["config_yaml"]))
    sim = FireSimulation(Config(config_dict=cfg))
    H, W = sim.config.area.screen_size
    pts = [(i, H // 4 - i) for i in range(H // 4 + 1)]  # skimage.draw.line(H // 4, 0, 0, W // 4) as (x, y)
    for kind in (3, 4, 5):
        sim.reset()
        sim.update_mitigation([(x, y, kind) for x, y in pts])
        got = [(int(x), int(y)) for y, x in np.argwhere(sim.fire_map == kind)]
        assert sorted(got) == sorted(pts)
    sim.reset()
    fire_map, _ = sim.run("1h")
    assert fire_map.max() == 2 and fire_map.dtype == np.int64
    sim.reset()
    sim.run(1)
    assert sim.elapsed_time == sim.config.simulation.update_rate and sim.elapsed_steps == 1
    sim.close()


def test_prune_after_max_fire_duration():
    """fire.py:116-161 timeline: the initial fire turns BURNED at the start of update max_dur + 1."""
    mgr = _simple_manager(pixel_scale=1e9, max_fire_duration=3)  # nothing else ever ignites
    fm = np.zeros((11, 11), dtype=np.int64)
    fm[5, 5] = 1
    for step in range(1, 4):
        fm, status = mgr.update(fm)
        assert fm[5, 5] == 1 and int(status) == 1, step
        assert mgr.durations == [step]
    fm, status = mgr.update(fm)  # sprite pruned, none left -> QUIT
    assert fm[5, 5] == 2 and int(status) == 0
    assert mgr.sprites == []
    mgr.close()


def test_attenuation_on_and_off():
    """test_fire.py:124-162: with attenuation every control line loses 980/490/245 per step,
    without it the lines' rate of spread is exactly zero."""
    lines = {(1, 1): 3, (2, 1): 4, (3, 1): 5, (6, 5): 3}  # (x, y) -> kind; (6, 5) touches the fire
    for attenuate in (True, False):
        mgr = _simple_manager(pixel_scale=50.0, attenuate_line_ros=attenuate, keep_rate_of_spread=True)
        fm = np.zeros((11, 11), dtype=np.int64)
        fm[5, 5] = 1
        for (x, y), k in lines.items():
            fm[y, x] = k
        fm, _ = mgr.update(fm)
        ros = mgr.rate_of_spread
        if attenuate:
            assert ros[1, 1] == -980 and ros[1, 2] == -490 and ros[1, 3] == -245
            assert ros[5, 6] == pytest.approx(mgr.rate_of_spread[5, 6]) and ros[5, 6] > -980
        else:
            assert ros[1, 1] == 0 and ros[1, 2] == 0 and ros[1, 3] == 0 and ros[5, 6] == 0
        assert np.array_equal(mgr.burn_amounts, ros)  # first step: burn == ros
        mgr.close()


def test_manager_errors_match_reference():
    with pytest.raises(ValueError, match="should match the terrain shape"):
        _simple_manager(U=np.zeros((3, 3)))
    with pytest.raises(ValueError, match="should be one of"):
        _simple_manager(U=[1.0, 2.0])
    mgr = _simple_manager()
    with pytest.raises(AssertionError):
        mgr.update(np.zeros((4, 4), dtype=np.int64))
    mgr.close()


@pytest.mark.parametrize("name", ["flat64", "gauss48"])
def test_fire_simulation_replays_reference_call_sequence(name):
    from simfire_b200.config import Config
    from simfire_b200.simulation import FireSimulation

    z = np.load(f"{GOLDEN}/api_sequence_{name}.npz")
    sim = FireSimulation(Config(config_dict=yaml.safe_load(str(z["config_yaml"]))))
    script = ast.literal_eval(str(z["script"]))
    k = 0
    for op, arg in script:
        if op == "run":
            fm, active = sim.run(arg)
            assert fm.dtype == np.int64 and fm.shape == z["maps"][k].shape
            assert np.array_equal(fm, z["maps"][k]), f"{name}: fire_map after call {k} ({op} {arg})"
            et, es, act = z["meta"][k]
            assert sim.elapsed_time == et and sim.elapsed_steps == int(es) and bool(active) == bool(act), k
            assert sim.active == bool(act)
            k += 1
        elif op == "mitigate":
            sim.update_mitigation(arg)
        elif op == "agents":
            sim.update_agent_positions(arg)
        elif op == "reset":
            sim.reset()
    assert np.array_equal(sim.agent_positions, z["agent_positions"])
    attr = sim.get_attribute_data()
    for key in ("w_0", "sigma", "delta", "M_x", "wind_speed", "wind_direction"):
        assert attr[key].dtype == z["attr_" + key].dtype, key
        assert np.array_equal(attr[key], z["attr_" + key]), key
    assert np.array_equal(np.asarray(attr["elevation"], dtype=np.float64).reshape(z["attr_elevation"].shape[:2]),
                          z["attr_elevation"].reshape(z["attr_elevation"].shape[:2]))  # fmt: skip
    assert sim.get_actions() == {"fireline": 3, "scratchline": 4, "wetline": 5}
    assert set(sim.get_attribute_bounds()) == set(sim.supported_attributes())
    sim.close()


def test_simulation_run_until_burned_out():
    """test_simulation.py:84-121: run('1h') on a 9x9 flat map ends with BURNED cells;
    run(1) advances elapsed_time by update_rate."""
    from simfire_b200.config import Config
    from simfire_b200.simulation import FireSimulation

    z = np.load(f"{GOLDEN}/api_sequence_flat64.npz")
    y = yaml.safe_load(str(z["config_yaml"]))
    y["area"]["screen_size"] = [9, 9]
    y["fire"]["fire_initial_position"]["static"]["position"] = "(4, 4)"
    sim = FireSimulation(Config(config_dict=y))
    fm, active = sim.run(1)
    assert sim.elapsed_time == y["simulation"]["update_rate"] and active
    fm, active = sim.run("1h")
    assert fm.max() == 2 and not active
    sim.close()


def test_save_data_history_matches_reference(tmp_path):
    """simulation.py:887-959 with `save_data: true`: the files written under
    <sf_home>/data/<start_time>/ -- names, metadata.json, static layers, the int8 fire_map history
    appended after every update -- against what the unmodified reference wrote for the same
    config and calls (tests/golden/gen_savedata_golden.py); then the JSON-lines variant."""
    import json

    from simfire_b200.config import Config
    from simfire_b200.simulation import FireSimulation

    z = np.load(f"{GOLDEN}/savedata_npy.npz")
    cfg = yaml.safe_load(str(z["config_yaml"]))
    cfg["simulation"]["sf_home"] = str(tmp_path / "home")
    sim = FireSimulation(Config(config_dict=cfg))
    sim.run(3)
    sim.update_mitigation([(x, 30, 3) for x in range(5, 40)])
    sim.run(4)
    datapath = tmp_path / "home" / "data" / sim.start_time
    assert sorted(p.name for p in datapath.iterdir()) == list(z["files"])
    history = np.load(datapath / "fire_map.npy")
    assert history.dtype == np.int8 and np.array_equal(history, z["fire_map"])
    want, got = json.loads(str(z["metadata_json"])), json.load(open(datapath / "metadata.json"))
    assert sorted(got) == sorted(want)
    for key in ("fire_map", "layer_types", "seeds", "shape", "static_data"):
        assert got[key] == want[key], key
    for name in ("w_0", "sigma", "delta", "M_x", "elevation", "wind_speed", "wind_direction"):
        a, b = np.load(datapath / f"{name}.npy"), z[f"static_{name}"]
        assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), name
    sim.close()

    cfg["simulation"]["data_type"] = "jsonl"
    cfg["simulation"]["sf_home"] = str(tmp_path / "home2")
    sim = FireSimulation(Config(config_dict=cfg))
    sim.run(2)
    lines = open(tmp_path / "home2" / "data" / sim.start_time / "fire_map.jsonl").read().splitlines()
    assert [list(json.loads(ln))[0] for ln in lines] == ["1", "2"]
    assert np.array_equal(np.array(json.loads(lines[1])["2"]), z["fire_map"][1])
    assert json.load(open(tmp_path / "home2" / "data" / sim.start_time / "w_0.json"))["data"][0][0] == pytest.approx(float(z["static_w_0"][0, 0]))
    sim.close()
    cfg["simulation"]["data_type"] = "csv"
    sim = FireSimulation(Config(config_dict=cfg))
    with pytest.raises(ValueError, match="Invalid data type"):
        sim.run(1)
    sim.close()


def test_load_mitigation_and_device_view():
    import torch

    from simfire_b200.config import Config
    from simfire_b200.simulation import FireSimulation

    z = np.load(f"{GOLDEN}/api_sequence_flat64.npz")
    sim = FireSimulation(Config(config_dict=yaml.safe_load(str(z["config_yaml"]))))
    m = np.zeros((64, 64), dtype=np.int64)
    m[10, :] = 3
    with pytest.warns(UserWarning, match="overwriting"):
        sim.load_mitigation(m)
    assert np.array_equal(sim.fire_map, m)
    bad = m.copy()
    bad[0, 0] = 17
    with pytest.warns(UserWarning, match="Invalid values"):
        sim.load_mitigation(bad)
    assert np.array_equal(sim.fire_map, m)
    t = sim.fire_map_device()
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), m.astype(np.int8))
    assert torch.count_nonzero(t == 3).item() == 64
    sim.close()


@pytest.mark.parametrize("name,starts", [("flat64", [(20, 24), (40, 10), (5, 50)]),
                                         ("gauss48", [(20, 24), (40, 10), (5, 40)])])
def test_batched_simulation_matches_single_envs(name, starts):
    """Also guards the slope planes of the batched class (gauss48 is not flat)."""
    from simfire_b200.config import Config
    from simfire_b200.simulation import BatchedFireSimulation, FireSimulation

    z = np.load(f"{GOLDEN}/api_sequence_{name}.npz")
    y = yaml.safe_load(str(z["config_yaml"]))
    W = y["area"]["screen_size"][1]
    batch = BatchedFireSimulation(Config(config_dict=y), 3, initial_positions=starts)
    batch.update_mitigation([(e, x, 30, 3) for e in range(3) for x in range(0, W)] + [(1, 2, 2, 9)])
    maps, active = batch.run(25)
    for e, pos in enumerate(starts):
        cfg = Config(config_dict=yaml.safe_load(str(z["config_yaml"])))
        cfg.reset_fire(pos=pos)
        single = FireSimulation(cfg)
        single.update_mitigation([(x, 30, 3) for x in range(0, W)])
        fm, act = single.run(25)
        assert np.array_equal(maps[e], fm), e
        assert bool(active[e]) == act
        assert batch.elapsed_time[e] == single.elapsed_time and batch.elapsed_steps[e] == single.elapsed_steps
        single.close()
    batch.close()


def test_seed_and_layer_type_wrappers_match_reference():
    """get_seeds / set_seeds / get_layer_types / set_fire_initial_position (simulation.py:574-829)
    against what the reference FireSimulation returned (tests/golden/gen_config_golden.py)."""
    import copy
    import json

    from simfire_b200.config import Config
    from simfire_b200.simulation import FireSimulation

    with open(f"{GOLDEN}/config_resets.json") as f:
        g = json.load(f)
    sim = FireSimulation(Config(config_dict=copy.deepcopy(g["config_dict"])))
    assert sim.get_seeds() == g["sim_get_seeds"]
    assert sim.get_layer_types() == g["sim_layer_types"]
    assert sim.set_seeds({"fuel": 42}) is g["sim_set_seeds_fuel"]
    assert sim.get_seeds() == g["sim_get_seeds_after"]
    with pytest.warns(UserWarning, match="No valid keys"):
        assert sim.set_seeds({"fuel": 43, "bogus": 1}) is g["sim_set_seeds_bad"]
    assert sim.get_seeds() == g["sim_get_seeds_after_bad"]  # the valid key was applied all the same
    with pytest.raises(KeyError):
        sim.set_seeds({"elevation": 5})
    with pytest.warns(UserWarning, match="No valid keys"):
        assert sim.set_layer_types({"bogus": "functional"}) is False
    assert sim.set_layer_types({"elevation": "functional", "fuel": "functional"}) is True

    def bbox(fm):
        ys, xs = np.nonzero(fm == 1)
        return [int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())]

    sim.set_fire_initial_position((40, 41))
    fm, _ = sim.run(1)  # config changes wait for the next reset()
    assert bbox(fm) == g["burning_bbox_before_reset"]
    sim.reset()
    fm, _ = sim.run(1)
    assert [sim.elapsed_time, sim.elapsed_steps] == g["elapsed_after_reset_run1"]
    attr = sim.get_attribute_data()
    assert [float(attr["w_0"][0, 0]), int(attr["sigma"][0, 0])] == g["fuel_after_reset"]
    assert bbox(fm) == g["burning_bbox_after_move"]
    sim.close()


def test_device_slopes_match_numpy_gradient():
    """sfb_set_elevation vs RothermelFireManager._compute_slopes (np.gradient in float64): after
    the float32 cast the planes agree (atan2 may differ in the last float64 bit, never more than
    one float32 ulp), and a scenario driven by device-computed slopes reproduces the reference."""
    from scenario_io import check_trajectory
    from test_gpu_parity import EngineAdapter, _burn_tol, engine_for

    from simfire_b200.workloads import compute_slopes

    sc = load_scenario("scenario_c_random_fuel_hills")
    planes = dict(sc["planes"])
    mag, ang = compute_slopes(sc["elevations"], float(sc["ps"]))
    assert np.array_equal(mag, planes["slope_mag"]) and np.array_equal(ang, planes["slope_dir"])
    planes["slope_mag"] = 0.0
    planes["slope_dir"] = 0.0
    with engine_for(sc) as eng:
        eng.set_static(planes)
        eng.set_elevation(sc["elevations"])
        eng.reset([sc["init"]])
        check_trajectory(sc, EngineAdapter(eng), **_burn_tol(sc))


def test_constant_spread_manager_replays_reference():
    """`ConstantSpreadFireManager.update` (fire.py:754-787) against trajectories recorded from the unmodified
    reference (tests/golden/gen_constant_golden.py): the reference's own test geometry (test_fire.py:399-470),
    a corner ignition, control lines / burned cells around the fire, a sprite pruned before it can spread,
    rate_of_spread 0 -- fire_map after every call, and how many sprites / durations are left."""
    from simfire_b200.enums import BurnStatus
    from simfire_b200.fire_manager import ConstantSpreadFireManager

    z = np.load(f"{GOLDEN}/constant_spread.npz")
    for name in z["names"]:
        H, W, x0, y0, max_dur, ros, calls = (int(v) for v in z[f"{name}_cfg"])
        mgr = ConstantSpreadFireManager((x0, y0), 2, max_dur, ros)
        fire_map = np.zeros((H, W))  # float64, like the reference's test
        for x, y, st in z[f"{name}_painted"]:
            fire_map[y, x] = st
        for k in range(calls):
            out = mgr.update(fire_map)
            assert out is fire_map and out.dtype == np.float64
            assert np.array_equal(out.astype(np.int8), z[f"{name}_maps"][k]), f"{name}: call {k + 1}"
            # sprites with a duration entry (the ones the reference's next prune keeps)
            assert len(mgr.durations) == min(int(z[f"{name}_n_durations"][k]), int(z[f"{name}_n_sprites"][k])), f"{name}: call {k + 1}"
        mgr.close()
    # the reference's own assertions (test_fire.py:431-470): nothing spreads before rate_of_spread calls, then all 8 neighbours
    mgr = ConstantSpreadFireManager((5, 2), 2, 4, 3)
    fm = np.zeros((20, 28))
    for _ in range(3):
        fm = mgr.update(fm)
        assert (fm == BurnStatus.BURNING).sum() == 0
    fm = mgr.update(fm)
    ys, xs = np.nonzero(fm == BurnStatus.BURNING)
    assert sorted(zip(xs.tolist(), ys.tolist())) == sorted((5 + dx, 2 + dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0))
    mgr.close()


def test_device_view_of_attributes_and_agent_positions():
    """`get_attribute_data` and `agent_positions` as device tensors (torch CUDA context
    needed: not run under the emulator).  The fuel / wind planes are zero-copy strided views of the stepper's
    own static records and must equal the host dictionaries of simulation.py:376-403."""
    import torch

    from simfire_b200.config import Config
    from simfire_b200.simulation import BatchedFireSimulation, FireSimulation

    z = np.load(f"{GOLDEN}/api_sequence_gauss48.npz")
    cfg = Config(config_dict=yaml.safe_load(str(z["config_yaml"])))
    sim = FireSimulation(cfg)
    host, dev = sim.get_attribute_data(), sim.get_attribute_data_device()
    assert set(dev) == set(host)
    for k, v in dev.items():
        assert v.is_cuda and tuple(v.shape) == tuple(cfg.area.screen_size)
        np.testing.assert_allclose(v.cpu().numpy().astype(np.float64), np.broadcast_to(np.asarray(host[k], dtype=np.float64), v.shape),
                                   rtol=1e-6 if k != "sigma" else 0)  # fmt: skip
    assert dev["sigma"].dtype == torch.int32 and dev["w_0"].dtype == torch.float32
    sim.close()
    bat = BatchedFireSimulation(cfg, 3)
    a = bat.agent_positions_device
    assert a.is_cuda and tuple(a.shape) == (3, *cfg.area.screen_size) and int(a.abs().sum()) == 0
    bat.update_agent_positions([(0, 5, 6, 1), (2, 7, 8, 2)])
    bat.update_agent_positions([(0, 9, 6, 1)])  # agent 1 of env 0 moves
    assert np.array_equal(bat.agent_positions_device.cpu().numpy(), bat.agent_positions.astype(np.int16))
    assert bat.agent_positions[0, 6, 9] == 1 and bat.agent_positions[0, 6, 5] == 0
    bat.reset(envs=[2])
    assert int(bat.agent_positions_device[2].abs().sum()) == 0
    for k, v in bat.get_attribute_data_device().items():
        assert v.is_cuda
    bat.close()


def test_reset_rebuilds_the_engine_when_the_config_changed():
    """The reference builds a new RothermelFireManager from the CURRENT config on every reset()
    (simulation.py:273-290) and its config objects are mutable; so must reset() here.  Also: `fire_map` is
    a read-only copy (in-place writes raise instead of being silently lost)."""
    from simfire_b200.config import Config
    from simfire_b200.simulation import FireSimulation

    z = np.load(f"{GOLDEN}/api_sequence_flat64.npz")
    sim = FireSimulation(Config(config_dict=yaml.safe_load(str(z["config_yaml"]))))
    sim.run(6)
    base = sim.fire_map.copy()
    with pytest.raises(ValueError):
        sim.fire_map[0, 0] = 3
    sim.config.fire.max_fire_duration = 1  # sprites burn out after one update
    sim.reset()
    sim.run(6)
    short = sim.fire_map.copy()
    assert not np.array_equal(short, base)
    ref = FireSimulation(sim.config)  # a fresh simulation of the edited config
    ref.run(6)
    assert np.array_equal(ref.fire_map, short)
    ref.close()
    sim.close()


def test_api_suite_on_the_bitboard_step():
    """The drop-in surface above (managers, FireSimulation replays of reference-recorded call sequences,
    save_data, spread graph, ConstantSpread) once more with every engine forced onto the bitboard step
    (SFB_FRONT=bits; the grids of these tests are small, so they get the dense sweep by default)."""
    import os
    import subprocess
    import sys

    if os.environ.get("SFB_FRONT"):
        pytest.skip("already running under SFB_FRONT")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SFB_FRONT="bits")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "tests/test_gpu_api.py",
           "tests/test_spread_graph.py", "-k", "not api_suite_on_the_bitboard_step" +
           (" and not device_view" if os.environ.get("SFB_EMULATED") else "")]  # fmt: skip
    res = subprocess.run(cmd, cwd=root, env=env, capture_output=True, text=True, timeout=1200)
    assert res.returncode == 0, "\n".join((res.stdout + res.stderr).splitlines()[-25:])
    assert " passed" in res.stdout
