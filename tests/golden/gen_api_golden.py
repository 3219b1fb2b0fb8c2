"""
Golden vectors for the FireSimulation-level API: runs the UNMODIFIED reference
`simfire.sim.simulation.FireSimulation` (headless, under ref_shim's stand-ins) through a
SimHarness-like call sequence modelled on the reference's `tests/sim.py`, and records what
every call returned.  Dev container only (needs /root/reference).

    python tests/golden/gen_api_golden.py   # rewrites tests/golden/api_sequence_*.npz
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
os.environ.setdefault("LOGLEVEL", "ERROR")
import ref_shim  # noqa: E402

ref_shim.install()
from simfire.sim.simulation import FireSimulation  # noqa: E402
from simfire.utils.config import Config  # noqa: E402

BASE = yaml.safe_load(open(os.path.join(ref_shim.REFERENCE_ROOT, "configs", "functional_config.yml")))


def make_config_dict(size, topo, start, wind=(7, 90.0), max_dur=4, diagonal=True, attenuate=True, ps=50,
                     runtime="24h", dt=1):
    y = copy.deepcopy(BASE)
    y["area"]["screen_size"] = [size[0], size[1]]
    y["area"]["pixel_scale"] = ps
    y["simulation"]["headless"] = True
    y["simulation"]["runtime"] = runtime
    y["simulation"]["update_rate"] = dt
    y["simulation"]["sf_home"] = "/tmp/sf_home_golden"
    y["mitigation"]["ros_attenuation"] = attenuate
    y["terrain"]["topography"]["functional"]["function"] = topo
    y["fire"]["fire_initial_position"]["static"]["position"] = f"({start[0]}, {start[1]})"
    y["fire"]["max_fire_duration"] = max_dur
    y["fire"]["diagonal_spread"] = diagonal
    y["wind"]["simple"]["speed"] = wind[0]
    y["wind"]["simple"]["direction"] = wind[1]
    return y


def record(name, cfg_dict, script):
    sim = FireSimulation(Config(config_dict=copy.deepcopy(cfg_dict)))
    maps, meta = [], []
    for op, arg in script:
        if op == "run":
            fm, active = sim.run(arg)
            maps.append(np.asarray(fm).astype(np.int8).copy())
            meta.append((float(sim.elapsed_time), int(sim.elapsed_steps), int(bool(active))))
        elif op == "mitigate":
            sim.update_mitigation(arg)
        elif op == "agents":
            sim.update_agent_positions(arg)
        elif op == "reset":
            sim.reset()
    attr = sim.get_attribute_data()
    np.savez_compressed(
        os.path.join(HERE, f"api_sequence_{name}.npz"),
        config_yaml=yaml.safe_dump(cfg_dict), script=repr(script), maps=np.stack(maps),
        meta=np.array(meta, dtype=np.float64), agent_positions=np.asarray(sim.agent_positions),
        attr_w_0=attr["w_0"], attr_sigma=attr["sigma"], attr_delta=attr["delta"], attr_M_x=attr["M_x"],
        attr_elevation=np.asarray(attr["elevation"], dtype=np.float64),
        attr_wind_speed=attr["wind_speed"], attr_wind_direction=attr["wind_direction"],
    )
    print(name, "calls", len(maps), "final burned", int((maps[-1] == 2).sum()), "meta", meta[-1])


def main():
    line = [(x, 40, 3) for x in range(10, 60)] + [(x, 41, 4) for x in range(10, 30)] + [(12, 12, 5), (12, 12, 3)]
    script = [("run", 1), ("run", 5), ("mitigate", line), ("agents", [(5, 6, 1), (7, 8, 2)]), ("run", 10),
              ("agents", [(9, 9, 1)]), ("run", "30m"), ("mitigate", [(30, 30, 5), (31, 30, 9)]), ("run", 3),
              ("reset", None), ("run", 4)]
    record("flat64", make_config_dict((64, 64), "flat", (20, 24)), script)
    # gaussian hill, 4-neighbour, no attenuation, slower wind, dt 2: exercises slopes and run() by time
    script2 = [("run", 2), ("mitigate", [(x, 20, 3) for x in range(0, 48)]), ("run", "40m"), ("run", 500)]
    record("gauss48", make_config_dict((48, 48), "gaussian", (25, 30), wind=(3, 200.0), max_dur=3, diagonal=False,
                                       attenuate=False, ps=30, runtime="3h", dt=2), script2)


if __name__ == "__main__":
    main()
