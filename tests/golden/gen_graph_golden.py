"""
Golden vectors for the fire-spread graph: edges of the reference's
`RothermelFireManager.fs_graph` (simfire/utils/graph.py) after running recorded scenarios
with the UNMODIFIED reference.  Dev container only.

    python tests/golden/gen_graph_golden.py   # rewrites tests/golden/graph_*.npz
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault("LOGLEVEL", "ERROR")
import ref_shim  # noqa: E402

fire_mod, roth_mod, enums, params_mod, presets = ref_shim.import_reference()
from scenario_io import load_scenario  # noqa: E402


def run(name, n_steps):
    sc = load_scenario(name)
    H, W = sc["H"], sc["W"]
    p = sc["planes"]
    fuels = np.empty((H, W), dtype=object)
    for y in range(H):
        for x in range(W):
            fuels[y, x] = params_mod.Fuel(p["w_0"][y, x], p["delta"][y, x], p["M_x"][y, x], p["sigma"][y, x])
    terrain = types.SimpleNamespace(fuels=fuels, elevations=sc["elevations"], screen_size=(H, W))
    env = params_mod.Environment(float(sc["M_f"]), p["U"], p["U_dir"])
    mgr = fire_mod.RothermelFireManager(
        sc["init"], 2, int(sc["max_dur"]), float(sc["ps"]), float(sc["dt"]), params_mod.FuelParticle(), terrain, env,
        max_time=sc["max_time"], attenuate_line_ros=bool(sc["attenuate"]), headless=True,
        diagonal_spread=bool(sc["diagonal"]))  # fmt: skip
    fire_map = np.full((H, W), enums.BurnStatus.UNBURNED)
    fire_map[sc["init"][1], sc["init"][0]] = enums.BurnStatus.BURNING
    for x, y, k in sc["pre"]:
        fire_map[y, x] = k
    for step in range(n_steps):
        fire_map, st = mgr.update(fire_map)
        if st != enums.GameStatus.RUNNING:
            break
    edges = np.array([(a[0], a[1], b[0], b[1]) for a, b in mgr.fs_graph.graph.edges], dtype=np.int32).reshape(-1, 4)
    np.savez_compressed(os.path.join(HERE, f"graph_{name}.npz"), edges=edges, n_steps=step + 1,
                        final_map=fire_map.astype(np.int8))
    print(name, "steps", step + 1, "edges", len(edges))


if __name__ == "__main__":
    run("scenario_a_models_diag_att", 60)
    run("scenario_b_models_4nbr_noatt", 60)
    run("scenario_c_random_fuel_hills", 40)
