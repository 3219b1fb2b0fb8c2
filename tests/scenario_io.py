"""Helpers shared by the parity tests: load golden scenarios, drive an implementation."""
from __future__ import annotations

import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PLANES = ("w_0", "delta", "M_x", "sigma", "U", "U_dir", "slope_mag", "slope_dir")


def scenario_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "scenario_*.npz")))


def load_scenario(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    sc = {k: z[k] for k in z.files}
    sc["name"] = name
    sc["H"], sc["W"] = int(sc["H"]), int(sc["W"])
    sc["planes"] = {k: sc["plane_" + k] for k in PLANES}
    sc["init"] = (int(sc["init"][0]), int(sc["init"][1]))
    sc["max_time"] = None if float(sc["max_time"]) < 0 else float(sc["max_time"])
    sched = {}
    for s, x, y, kind in sc["sched_points"]:
        sched.setdefault(int(s), []).append((int(x), int(y), int(kind)))
    sc["schedule"] = sched
    sc["pre"] = [(int(x), int(y), int(k)) for x, y, k in sc["pre_points"]]
    return sc


def dense_params(sc):
    from oracle.dense_numpy import DenseParams

    return DenseParams(
        pixel_scale=float(sc["ps"]), update_rate=float(sc["dt"]), max_fire_duration=int(sc["max_dur"]),
        max_time=sc["max_time"], attenuate_line_ros=bool(sc["attenuate"]),
        diagonal_spread=bool(sc["diagonal"]), M_f=float(sc["M_f"]),
    )  # fmt: skip


def check_trajectory(sc, sim, *, burn_exact=True, burn_rtol=0.0, burn_atol=0.0):
    """
    Drive ``sim`` (anything with apply_points / step / status-map / burn / elapsed accessors)
    through the scenario and compare with the recorded reference trajectory.

    sim API: apply_points(points), step() -> int status, get_map() -> int8 (H, W),
             get_burn() -> f64 (H, W), get_ros() -> f64 (H, W), elapsed() -> float
    """
    map_at = {int(s): i for i, s in enumerate(sc["map_steps"])}
    burn_at = {int(s): i for i, s in enumerate(sc["burn_steps"])}
    sim.apply_points(sc["pre"])
    n = int(sc["n_steps"])
    for step in range(1, n + 1):
        pts = sc["schedule"].get(step)
        if pts:
            sim.apply_points(pts)
        st = sim.step()
        assert st == int(sc["status"][step - 1]), f"{sc['name']}: status at step {step}"
        assert sim.elapsed() == float(sc["elapsed"][step - 1]), f"{sc['name']}: elapsed at step {step}"
        if step in map_at:
            got = sim.get_map()
            want = sc["maps"][map_at[step]]
            if not np.array_equal(got, want):
                bad = np.argwhere(got != want)
                raise AssertionError(
                    f"{sc['name']}: fire_map differs at step {step} in {len(bad)} cells, first {bad[:5].tolist()}"
                )
        if step in burn_at:
            want = sc["burns"][burn_at[step]]
            got = sim.get_burn()
            if burn_exact:
                assert np.array_equal(got, want), f"{sc['name']}: burn differs at step {step}"
            else:
                np.testing.assert_allclose(got, want, rtol=burn_rtol, atol=burn_atol,
                                           err_msg=f"{sc['name']}: burn at step {step}")  # fmt: skip
            if int(sc["status"][step - 1]) == 1 and hasattr(sim, "get_ros"):
                ros = sim.get_ros()
                if ros is not None:
                    want = sc["ross"][burn_at[step]]
                    if burn_exact:
                        assert np.array_equal(ros, want), f"{sc['name']}: ros differs at step {step}"
                    else:
                        np.testing.assert_allclose(ros, want, rtol=burn_rtol, atol=burn_atol)
