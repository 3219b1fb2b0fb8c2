"""
ORACLE (test infrastructure, not product code) -- NumPy restatement of the reference's
Rothermel rate-of-spread arithmetic.

Follows ``/root/reference/simfire/world/rothermel.py:4-136`` operation by operation,
with the dtypes the reference actually produces when it is called from
``RothermelFireManager.update`` (``simfire/game/managers/fire.py:675-693``) with the 17
float32 vectors built by ``_flatten_params`` (``fire.py:519-548``):

* every fuel / wind term is evaluated in float32 (Python scalars are "weak", the arrays
  are float32);
* ``sign`` (``rothermel.py:118``) is int64, so ``phi_s`` (``:119``) and everything after
  it is float64;
* ``x ** 2`` is ``np.square`` and ``x ** 0.5`` is ``np.sqrt`` (NumPy's scalar-power fast
  path); every other ``**`` is ``np.power`` in float32.

Because it issues the same NumPy ufuncs in the same order it is bit-identical to the
reference on the same NumPy build (pinned in ``tests/test_oracle_numpy.py`` against
``tests/golden/*.npz`` generated from the reference itself by
``tests/golden/gen_golden.py``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu-baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32

# Offsets dst - src in the reference's neighbour order (fire.py:211-221) as (dx, dy),
# image coordinates (y grows downwards).
NEIGHBOUR_OFFSETS = (
    (1, 0),
    (1, 1),
    (0, 1),
    (-1, 1),
    (-1, 0),
    (-1, -1),
    (0, -1),
    (1, -1),
)
# 4-neighbour variant (fire.py:223-228) uses directions 0, 2, 4, 6 of the table above.
NEIGHBOUR_DIRS_4 = (0, 2, 4, 6)


def travel_angles() -> np.ndarray:
    """theta = arctan2(y_src - y_dst, x_dst - x_src) in float32 (rothermel.py:102)."""
    dx = np.array([o[0] for o in NEIGHBOUR_OFFSETS], dtype=f32)
    dy = np.array([o[1] for o in NEIGHBOUR_OFFSETS], dtype=f32)
    return np.arctan2(f32(0) - dy, dx)  # 0 - dy, not -dy: keeps +0.0 so west is +pi


def rate_of_spread(
    direction: np.ndarray,
    w_0: np.ndarray,
    delta: np.ndarray,
    M_x: np.ndarray,
    sigma: np.ndarray,
    U: np.ndarray,
    U_dir: np.ndarray,
    slope_mag: np.ndarray,
    slope_dir: np.ndarray,
    *,
    h: float = 8000.0,
    S_T: float = 0.0555,
    S_e: float = 0.01,
    p_p: float = 32.0,
    M_f: float = 0.03,
    theta: np.ndarray | None = None,
) -> np.ndarray:
    """
    Rate of spread (ft/min, float64) for N (src -> dst) pairs.

    ``direction[i]`` indexes NEIGHBOUR_OFFSETS (dst = src + offset); all other arrays are
    the DESTINATION cell's values (fire.py:481-497) and are cast to float32 here exactly
    as ``_flatten_params`` does.  ``theta`` overrides the travel angle (used to replay the
    reference unit test, which passes src == dst so theta = 0).
    """
    direction = np.asarray(direction)
    n = direction.shape[0]
    cast = lambda a: np.broadcast_to(np.asarray(a, dtype=np.float64).astype(f32), (n,))  # noqa: E731
    w_0, delta, M_x, sigma = cast(w_0), cast(delta), cast(M_x), cast(sigma)
    U, U_dir, slope_mag, slope_dir = cast(U), cast(U_dir), cast(slope_mag), cast(slope_dir)
    h_, S_T_, S_e_, p_p_, M_f_ = cast(h), cast(S_T), cast(S_e), cast(p_p), cast(M_f)
    if theta is None:
        theta = travel_angles()[direction]
    theta = np.asarray(theta, dtype=f32)

    out = np.zeros(n, dtype=np.float64)
    keep = np.nonzero(w_0 > 0)[0]  # rothermel.py:54 -- non-burnable pairs stay 0
    if keep.size == 0:
        return out
    w_0, delta, M_x, sigma = w_0[keep], delta[keep], M_x[keep], sigma[keep]
    U, U_dir, slope_mag, slope_dir = U[keep], U_dir[keep], slope_mag[keep], slope_dir[keep]
    h_, S_T_, S_e_, p_p_, M_f_ = h_[keep], S_T_[keep], S_e_[keep], p_p_[keep], M_f_[keep]
    theta = theta[keep]
    one = np.ones_like(w_0)

    with np.errstate(all="ignore"):
        # --- fuel-only terms, float32 (rothermel.py:74-98)
        eta_S = np.minimum(0.174 * S_e_**-0.19, one)
        r_M = np.minimum(M_f_ / M_x, one)
        eta_M = 1 - 2.59 * r_M + 5.11 * r_M**2 - 3.52 * r_M**3
        w_n = w_0 * (1 - S_T_)
        p_b = w_0 / delta
        B = p_b / p_p_
        B_op = 3.348 * sigma**-0.8189
        s15 = sigma**1.5
        gamma_max = s15 / (495 + 0.0594 * sigma**1.5)
        A = 133 * sigma**-0.7913
        ratio = B / B_op
        gamma = gamma_max * ratio**A * np.exp(A * (1 - ratio))
        I_R = gamma * w_n * h_ * eta_M * eta_S
        xi = np.exp((0.792 + 0.681 * sigma**0.5) * (B + 0.1)) / (192 + 0.2595 * sigma)
        c = 7.47 * np.exp(-0.133 * sigma**0.55)
        b = 0.02526 * sigma**0.54
        e = 0.715 * np.exp(-3.59e-4 * sigma)
        # --- wind factor, float32 (rothermel.py:102-111)
        psi = np.radians(90 - U_dir)
        U_along = np.maximum(U * np.cos(psi - theta), np.zeros_like(U))
        phi_w = c * U_along**b * ratio**-e
        # --- slope factor: float32 until multiplied by the int64 sign (rothermel.py:117-119)
        s_along = -slope_mag * np.cos(slope_dir + theta)
        sign = -1 + 2 * (s_along > 0)
        phi_s = 5.275 * B**-0.3 * sign * s_along**2
        # --- heat sink, float32 (rothermel.py:121-123)
        eps = np.exp(-138 / sigma)
        Q_ig = 250 + 1116 * M_f_
        # --- assemble in float64 (rothermel.py:128, :134)
        R = ((I_R * xi) * (1 + phi_w + phi_s)) / (p_b * eps * Q_ig)
    assert R.dtype == np.float64
    out[keep] = R
    return np.maximum(out, np.zeros_like(out))
