"""
Golden values for the episode-to-episode config mutation API: runs the UNMODIFIED reference
`Config.reset_fire / reset_terrain / reset_wind` and `FireSimulation.get_seeds / set_seeds /
get_layer_types / set_fire_initial_position` (headless, under ref_shim's stand-ins) and records
what they produce.  Dev container only (needs /root/reference).

    python tests/golden/gen_config_golden.py   # rewrites tests/golden/config_resets.json
"""
from __future__ import annotations

import copy
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
os.environ.setdefault("LOGLEVEL", "ERROR")
import ref_shim  # noqa: E402

ref_shim.install()
from gen_api_golden import make_config_dict  # noqa: E402
from simfire.sim.simulation import FireSimulation  # noqa: E402
from simfire.utils.config import Config  # noqa: E402


def fuel_of(cfg):
    f = cfg.terrain.fuel_layer.data[0, 0, 0]
    return [float(f.w_0), float(f.delta), float(f.M_x), float(f.sigma)]


def main():
    out = {}
    base = make_config_dict((64, 64), "flat", (20, 24))
    out["config_dict"] = base

    # static start: pos moves it, seed is ignored
    c = Config(config_dict=copy.deepcopy(base))
    c.reset_fire(pos=(3, 9))
    out["static_after_pos"] = [int(v) for v in c.fire.fire_initial_position]
    c.reset_fire(77)
    out["static_after_seed"] = [int(v) for v in c.fire.fire_initial_position]
    out["static_seed_attr"] = c.fire.seed

    # random start: seed re-draws it, pos is ignored
    y = copy.deepcopy(base)
    y["fire"]["fire_initial_position"]["type"] = "random"
    y["area"]["screen_size"] = [48, 48]  # the reference's fuel layer cannot build non-square screens
    c = Config(config_dict=y)
    out["random_initial"] = [int(v) for v in c.fire.fire_initial_position] + [c.fire.seed]
    draws = {}
    for seed in (0, 5, 99, 123456):
        c.reset_fire(seed)
        draws[str(seed)] = [int(v) for v in c.fire.fire_initial_position] + [c.fire.seed]
    out["random_draws_48x48"] = draws
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        c.reset_fire(pos=(1, 2))
    out["random_after_pos"] = [int(v) for v in c.fire.fire_initial_position]

    # fuel seed
    c = Config(config_dict=copy.deepcopy(base))
    out["fuel_initial"] = fuel_of(c)
    fuels = {}
    for seed in (0, 7, 1113):
        c.reset_terrain(fuel_seed=seed)
        fuels[str(seed)] = fuel_of(c)
    out["fuel_by_seed"] = fuels
    c.reset_wind(speed_seed=3, direction_seed=4)
    out["wind_after_reset"] = [float(c.wind.speed[0, 0]), float(c.wind.direction[0, 0])]

    # FireSimulation wrappers
    sim = FireSimulation(Config(config_dict=copy.deepcopy(base)))
    out["sim_get_seeds"] = sim.get_seeds()
    out["sim_layer_types"] = sim.get_layer_types()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["sim_set_seeds_fuel"] = bool(sim.set_seeds({"fuel": 42}))
        out["sim_get_seeds_after"] = sim.get_seeds()
        out["sim_set_seeds_bad"] = bool(sim.set_seeds({"fuel": 43, "bogus": 1}))
        out["sim_get_seeds_after_bad"] = sim.get_seeds()
        try:
            out["sim_set_seeds_elevation_flat"] = bool(sim.set_seeds({"elevation": 5}))
        except KeyError as e:  # 'flat' has no block under terrain.topography.functional
            out["sim_set_seeds_elevation_flat"] = f"KeyError({e})"
    sim.set_fire_initial_position((40, 41))
    fm_before, _ = sim.run(1)  # the move takes effect at the next reset() only
    ys, xs = np.nonzero(np.asarray(fm_before) == 1)
    out["burning_bbox_before_reset"] = [int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())]
    sim.reset()
    fm, _ = sim.run(1)
    out["elapsed_after_reset_run1"] = [float(sim.elapsed_time), int(sim.elapsed_steps)]
    out["fuel_after_reset"] = [float(sim.get_attribute_data()["w_0"][0, 0]), int(sim.get_attribute_data()["sigma"][0, 0])]
    ys, xs = np.nonzero(np.asarray(fm) == 1)
    out["burning_bbox_after_move"] = [int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())]
    with open(os.path.join(HERE, "config_resets.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps({k: v for k, v in out.items() if k != "config_dict"}, indent=1))


if __name__ == "__main__":
    main()
