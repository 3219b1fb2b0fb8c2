"""Fire-spread graph reconstruction (simfire_b200/graph.py) against the edges of the reference's
networkx graph (tests/golden/gen_graph_golden.py): on the CPU from the oracle's ignition plane,
on the GPU from the device's (`keep_ignition=True`)."""
import numpy as np
import pytest
from scenario_io import GOLDEN, dense_params, load_scenario

from simfire_b200.graph import edge_set, spread_edges, to_networkx

NAMES = ["scenario_a_models_diag_att", "scenario_b_models_4nbr_noatt", "scenario_c_random_fuel_hills"]


def _oracle_ignition(sc, n_steps):
    from oracle.dense_numpy import DenseFire

    o = DenseFire(sc["planes"], dense_params(sc), sc["init"])
    o.apply_points(sc["pre"])
    for _ in range(n_steps):
        if o.step() != 1:
            break
    return o


@pytest.mark.parametrize("name", NAMES)
def test_edges_from_oracle_ignition_plane_match_reference_graph(name):
    sc = load_scenario(name)
    z = np.load(f"{GOLDEN}/graph_{name}.npz")
    o = _oracle_ignition(sc, int(z["n_steps"]))
    assert np.array_equal(o.status, z["final_map"])
    got = edge_set(spread_edges(o.ign, int(sc["max_dur"])))
    want = edge_set(z["edges"])
    assert got == want, f"missing {list(want - got)[:5]} extra {list(got - want)[:5]}"


def test_networkx_view():
    sc = load_scenario(NAMES[0])
    z = np.load(f"{GOLDEN}/graph_{NAMES[0]}.npz")
    o = _oracle_ignition(sc, int(z["n_steps"]))
    g = to_networkx(o.ign, int(sc["max_dur"]))
    assert g.number_of_nodes() == sc["H"] * sc["W"] and g.number_of_edges() == len(z["edges"])
    assert g.in_degree(sc["init"]) == 0  # the initial fire has no parent


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_ignition_plane_and_graph(name):
    from simfire_b200 import FireEngine

    sc = load_scenario(name)
    z = np.load(f"{GOLDEN}/graph_{name}.npz")
    n = int(z["n_steps"])
    o = _oracle_ignition(sc, n)
    with FireEngine(sc["H"], sc["W"], 1, pixel_scale=float(sc["ps"]), update_rate=float(sc["dt"]),
                    max_fire_duration=int(sc["max_dur"]), max_time=sc["max_time"],
                    attenuate_line_ros=bool(sc["attenuate"]), diagonal_spread=bool(sc["diagonal"]),
                    M_f=float(sc["M_f"]), keep_ignition=True) as eng:  # fmt: skip
        eng.set_static(sc["planes"])
        eng.reset([sc["init"]])
        eng.apply_points([(0, x, y, k) for x, y, k in sc["pre"]])
        eng.step(n)
        ign = eng.plane("ignition")
        assert np.array_equal(ign, o.ign)
        assert edge_set(spread_edges(ign, int(sc["max_dur"]))) == edge_set(z["edges"])
        with pytest.raises(Exception):
            FireEngine(8, 8, 1, pixel_scale=1.0, update_rate=1.0, max_fire_duration=3).plane("ignition")
