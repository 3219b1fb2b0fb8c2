"""
The NumPy oracle (oracle/rothermel_numpy.py, oracle/dense_numpy.py) against golden vectors
recorded from the unmodified reference (tests/golden/gen_golden.py).

In the dev container (same NumPy build that produced the vectors) the match is bit-exact.
On another CPU / NumPy build NumPy's float32 SIMD pow/exp/cos may differ by an ulp, so R is
held to 2e-6 relative (the north-star tolerance is 1e-5) while fire_map stays exact: the
scenarios were selected with an ignition margin (see gen_golden.py).
"""
import numpy as np
import pytest
from scenario_io import GOLDEN, check_trajectory, dense_params, load_scenario, scenario_names

from oracle.dense_numpy import DenseFire
from oracle.rothermel_numpy import rate_of_spread, travel_angles


def test_travel_angle_table():
    # theta per neighbour direction, float32
    want = np.array([0, -0.7853981256484985, -1.5707963705062866, -2.356194496154785, 3.1415927410125732,
                     2.356194496154785, 1.5707963705062866, 0.7853981256484985], dtype=np.float32)  # fmt: skip
    assert np.array_equal(travel_angles(), want)


def test_rothermel_pairs_golden():
    z = np.load(f"{GOLDEN}/rothermel_pairs.npz")
    h, S_T, S_e, p_p, M_f = z["consts"]
    R = rate_of_spread(z["direction"], z["w_0"], z["delta"], z["M_x"], z["sigma"], z["U"], z["U_dir"],
                       z["slope_mag"], z["slope_dir"], h=h, S_T=S_T, S_e=S_e, p_p=p_p, M_f=M_f)  # fmt: skip
    assert R.dtype == np.float64
    assert np.array_equal(R == 0, z["R"] == 0)
    np.testing.assert_allclose(R, z["R"], rtol=2e-6, atol=0)


def test_reference_known_answer():
    """simfire/world/_tests/test_rothermel.py:10-100 (places=2)."""
    z = np.load(f"{GOLDEN}/rothermel_pairs.npz")
    R = rate_of_spread(np.zeros(8, int), z["kat_w_0"], z["kat_delta"], z["kat_M_x"], z["kat_sigma"], z["kat_U"],
                       z["kat_U_dir"], np.zeros(8), np.zeros(8), M_f=0.03, theta=np.zeros(8, np.float32))  # fmt: skip
    np.testing.assert_array_almost_equal(R, z["kat_literal"], decimal=2)
    np.testing.assert_allclose(R, z["kat_R"], rtol=2e-6)


class _Adapter:
    def __init__(self, sim):
        self.sim = sim

    def apply_points(self, pts):
        self.sim.apply_points(pts)

    def step(self):
        return self.sim.step()

    def get_map(self):
        return self.sim.status

    def get_burn(self):
        return self.sim.burn

    def get_ros(self):
        return self.sim.ros

    def elapsed(self):
        return self.sim.elapsed_time


@pytest.mark.parametrize("name", scenario_names())
def test_dense_oracle_matches_reference_trajectory(name):
    sc = load_scenario(name)
    sim = DenseFire(sc["planes"], dense_params(sc), sc["init"])
    check_trajectory(sc, _Adapter(sim), burn_exact=False, burn_rtol=2e-6, burn_atol=1e-9)
