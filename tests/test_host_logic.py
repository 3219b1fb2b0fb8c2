"""CPU tests of the host-side mirror: config parsing, unit helpers, workload generators."""
import numpy as np
import pytest
import yaml
from scenario_io import GOLDEN


def _cfg_dict():
    z = np.load(f"{GOLDEN}/api_sequence_flat64.npz")
    return yaml.safe_load(str(z["config_yaml"]))


def test_str_to_minutes():
    from simfire_b200.config import str_to_minutes

    assert str_to_minutes("24h") == 1440
    assert str_to_minutes("1h 30m") == 90
    assert str_to_minutes("2d") == 2880
    assert str_to_minutes("90") == 90


def test_chaparral_matches_reference_values():
    """chaparral(seed=1113) as produced by simfire/utils/terrain.py:93-114."""
    from simfire_b200.config import chaparral

    f = chaparral(1113)
    assert (f.w_0, f.delta, f.M_x, f.sigma) == (0.9810356625846572, 5.890006842991012, 0.9833113830744984,
                                                 3433.643783383716)  # fmt: skip


def test_config_sections_and_units():
    from simfire_b200.config import Config

    c = Config(config_dict=_cfg_dict())
    assert c.area.screen_size == (64, 64) and c.area.pixel_scale == 50
    assert c.simulation.runtime == 1440 and c.simulation.update_rate == 1
    assert c.fire.fire_initial_position == (20, 24) and c.fire.max_fire_duration == 4 and c.fire.diagonal_spread
    assert c.mitigation.ros_attenuation is True and c.environment.moisture == 0.03
    assert c.wind.speed.dtype == np.float64 and c.wind.speed.shape == (64, 64)
    assert np.all(c.wind.speed == 616.0) and np.all(c.wind.direction == 90.0)
    assert c.terrain.fuel_layer.data.shape == (64, 64, 1) and c.terrain.topography_layer.data.shape == (64, 64, 1)


def test_config_random_start_and_errors():
    from simfire_b200.config import Config, ConfigError

    y = _cfg_dict()
    y["fire"]["fire_initial_position"]["type"] = "random"
    y["fire"]["fire_initial_position"]["random"]["seed"] = 1234
    c = Config(config_dict=y)
    rng = np.random.default_rng(1234)
    assert c.fire.fire_initial_position == (int(rng.integers(64, dtype=int)), int(rng.integers(64, dtype=int)))
    y = _cfg_dict()
    y["wind"]["function"] = "perlin"
    with pytest.raises(ConfigError):
        Config(config_dict=y)
    y = _cfg_dict()
    y["terrain"]["topography"]["type"] = "operational"
    with pytest.raises(ConfigError):
        Config(config_dict=y)
    with pytest.raises(ConfigError):
        Config()


def test_config_from_arrays():
    from simfire_b200.config import Config

    fuels = np.zeros((8, 12, 4))
    c = Config.from_arrays(fuels=fuels, elevations=np.ones((8, 12)), wind_speed=5.0, wind_direction=np.full((8, 12), 10.0),
                           pixel_scale=98, fire_initial_position=(3, 4), runtime=100, moisture=0.001)  # fmt: skip
    assert c.area.screen_size == (8, 12) and c.simulation.runtime == 100 and c.wind.speed.shape == (8, 12)
    assert c.fire.fire_initial_position == (3, 4)


def test_workload_generators_are_deterministic():
    from simfire_b200.workloads import cfg1_functional_flat, synthetic_operational

    a, b = synthetic_operational(96, 160, seed=3), synthetic_operational(96, 160, seed=3)
    for k in a.planes:
        assert np.array_equal(a.planes[k], b.planes[k])
    assert a.planes["w_0"][a.init_pos[1], a.init_pos[0]] > 0
    assert a.pixel_scale == 98.0 and a.max_fire_duration == 5 and not a.attenuate_line_ros
    starts = a.burnable_starts(50, seed=1)
    assert np.all(a.planes["w_0"][starts[:, 1], starts[:, 0]] > 0)
    flat = synthetic_operational(64, 64, seed=0, flat=True)
    assert np.all(flat.planes["slope_mag"] == 0)
    c1 = cfg1_functional_flat()
    assert c1.H == 128 and c1.planes["U"][0, 0] == 616.0 and c1.max_time == 1440.0


def test_compute_slopes_matches_oracle_helper():
    from oracle.dense_numpy import compute_slopes as ref
    from simfire_b200.workloads import compute_slopes

    e = np.random.default_rng(0).uniform(0, 500, (17, 23))
    for a, b in zip(compute_slopes(e, 30.0), ref(e, 30.0)):
        assert np.array_equal(a, b)


def _resets():
    import json

    with open(f"{GOLDEN}/config_resets.json") as f:
        return json.load(f)


def _fuel_of(cfg):
    f = cfg.terrain.fuel_layer.data[0, 0, 0]
    return [f.w_0, f.delta, f.M_x, f.sigma]


def test_config_reset_fire_matches_reference():
    """Config.reset_fire (config.py:1088-1133) against values recorded from the reference."""
    import copy

    from simfire_b200.config import Config

    g = _resets()
    c = Config(config_dict=copy.deepcopy(g["config_dict"]))
    c.reset_fire(pos=(3, 9))
    assert list(c.fire.fire_initial_position) == g["static_after_pos"]
    with pytest.warns(UserWarning, match="does not support"):
        c.reset_fire(77)  # a seed means nothing to a static start
    assert list(c.fire.fire_initial_position) == g["static_after_seed"] and c.fire.seed == g["static_seed_attr"]
    with pytest.raises(ValueError):
        c.reset_fire()
    with pytest.raises(ValueError):
        c.reset_fire(1, (2, 3))

    y = copy.deepcopy(g["config_dict"])
    y["fire"]["fire_initial_position"]["type"] = "random"
    y["area"]["screen_size"] = [48, 48]
    c = Config(config_dict=y)
    assert list(c.fire.fire_initial_position) + [c.fire.seed] == g["random_initial"]
    for seed, want in sorted(g["random_draws_48x48"].items(), key=lambda kv: int(kv[0])):
        c.reset_fire(int(seed))
        assert list(c.fire.fire_initial_position) + [c.fire.seed] == want
    with pytest.warns(UserWarning):
        c.reset_fire(pos=(1, 2))
    assert list(c.fire.fire_initial_position) == g["random_after_pos"]


def test_config_reset_terrain_and_wind_match_reference():
    import copy

    from simfire_b200.config import Config, ConfigError

    g = _resets()
    c = Config(config_dict=copy.deepcopy(g["config_dict"]))
    assert _fuel_of(c) == g["fuel_initial"]
    for seed, want in g["fuel_by_seed"].items():
        c.reset_terrain(fuel_seed=int(seed))
        assert _fuel_of(c) == want
        assert c.terrain.fuel_function.kwargs["seed"] == int(seed)
    c.reset_wind(speed_seed=3, direction_seed=4)
    assert [c.wind.speed[0, 0], c.wind.direction[0, 0]] == g["wind_after_reset"]
    assert c.wind.speed_function is None and c.wind.direction_function is None
    with pytest.raises(KeyError):  # the reference fails the same way: 'flat' has no YAML block to hold a seed
        c.reset_terrain(topography_seed=5)
    with pytest.raises(ConfigError):
        c.reset_terrain(fuel_type="operational")
    arr = Config.from_arrays(fuels=np.zeros((8, 8, 4)), elevations=np.zeros((8, 8)), wind_speed=1.0, wind_direction=0.0,
                             pixel_scale=30, fire_initial_position=(1, 1))  # fmt: skip
    arr.reset_fire(pos=(5, 6))
    assert arr.fire.fire_initial_position == (5, 6)
    with pytest.raises(ConfigError):
        arr.reset_terrain(fuel_seed=1)
    with pytest.raises(ConfigError):
        arr.reset_wind()
