"""
Golden trajectories of the UNMODIFIED reference's `ConstantSpreadFireManager.update`
(simfire/game/managers/fire.py:722-787), run in the dev container under oracle/ref_shim.py:

    python tests/golden/gen_constant_golden.py      # rewrites tests/golden/constant_spread.npz

Scenarios: the reference's own test geometry (test_fire.py:399-470: rate_of_spread =
max_fire_duration - 1), a corner ignition, control lines and burned cells around the fire,
rate_of_spread >= max_fire_duration (pruned before it can spread), rate_of_spread = 0.  The
fire_map after every call, the sprite positions and the durations list are recorded.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

fire_mod, _, enums, _, _ = ref_shim.import_reference()

CASES = [
    # name, (H, W), init (x, y), max_fire_duration, rate_of_spread, calls, painted cells (x, y, status)
    ("reference_test", (20, 28), (5, 2), 4, 3, 8, []),
    ("corner", (9, 7), (0, 0), 3, 1, 6, []),
    ("lines_and_burned", (12, 12), (6, 6), 5, 2, 9, [(5, 5, 3), (6, 5, 4), (7, 5, 5), (7, 6, 2), (5, 7, 1), (6, 7, 3)]),
    ("pruned_first", (8, 8), (3, 3), 2, 2, 5, []),
    ("spreads_at_once", (8, 8), (7, 4), 3, 0, 5, [(6, 3, 5)]),
]

out = {}
for name, (H, W), init, max_dur, ros, calls, painted in CASES:
    mgr = fire_mod.ConstantSpreadFireManager(init, 2, max_dur, ros)
    fire_map = np.zeros((H, W))  # what the reference's test passes (float64, test_fire.py:422)
    for x, y, st in painted:
        fire_map[y, x] = st
    maps, sprites, durs = [], [], []
    for _ in range(calls):
        fire_map = mgr.update(fire_map)
        maps.append(fire_map.astype(np.int8).copy())
        sprites.append(sorted((int(s.rect.x), int(s.rect.y)) for s in mgr.sprites))
        durs.append(list(mgr.durations))
    out[f"{name}_cfg"] = np.array([H, W, init[0], init[1], max_dur, ros, calls], dtype=np.int32)
    out[f"{name}_painted"] = np.array(painted, dtype=np.int32).reshape(-1, 3)
    out[f"{name}_maps"] = np.stack(maps)
    out[f"{name}_n_sprites"] = np.array([len(s) for s in sprites], dtype=np.int32)
    out[f"{name}_n_durations"] = np.array([len(d) for d in durs], dtype=np.int32)
    print(name, [len(s) for s in sprites], durs[-1], int((maps[-1] == 1).sum()), int((maps[-1] == 2).sum()))
out["names"] = np.array([c[0] for c in CASES])
np.savez_compressed(os.path.join(HERE, "constant_spread.npz"), **out)
