"""Device WGS84 solvers (hhmarl_2d_b200/csrc/hh_geodesic.cuh) against the oracle's restatement of
geographiclib, over the map box of the env (lat 5..5.5, lon 7..7.5), plus the accuracy claim
behind the local inverse that the kernels use for threshold decisions."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev(mode, arr):
    from hhmarl_2d_b200 import _native as nat
    arr = np.ascontiguousarray(arr, np.float64)
    n = arr.shape[1]
    out = np.empty((2, n), np.float64)
    nat.check(nat.lib().hh_debug_geodesic(mode, n, arr.ctypes.data, out.ctypes.data), "hh_debug_geodesic")
    return out


def _pairs(rng, n, max_deg):
    lat1 = rng.uniform(4.99, 5.51, n); lon1 = rng.uniform(6.99, 7.51, n)
    ang = rng.uniform(0, 2 * np.pi, n); r = rng.uniform(0, max_deg, n)
    return np.stack([lat1, lon1, lat1 + r * np.cos(ang), lon1 + r * np.sin(ang)])


def test_direct_matches_oracle():
    import oracle as orc
    rng = np.random.default_rng(1)
    n = 20000
    x = np.stack([rng.uniform(4.99, 5.51, n), rng.uniform(6.99, 7.51, n), rng.uniform(0, 360, n),
                  rng.uniform(0, 1100, n)])
    x[2, :100] = np.repeat([0.0, 90.0, 180.0, 270.0, 359.0], 20)      # cardinal headings
    x[3, 100:110] = 0.0
    got = _dev(0, x)
    exp = np.array([orc.geod_direct(*c)[:2] for c in x.T]).T
    assert np.abs(got - exp).max() < 5e-14                                # degrees (~5 nm)


def test_direct_short_on_device_matches_oracle_and_quadrature():
    """geo::direct_short as the v4 step kernel calls it (mode 3: heading vector, mode 4: sincosd) on the device."""
    import oracle as orc
    from test_oracle_geodesic import _exact_direct
    rng = np.random.default_rng(7)
    n = 20000
    x = np.stack([rng.uniform(4.99, 5.51, n), rng.uniform(6.99, 7.51, n),
                  np.where(rng.random(n) < 0.5, rng.integers(0, 360, n).astype(float), rng.uniform(-20, 400, n)),
                  np.where(rng.random(n) < 0.2, rng.uniform(250, 1030, n), rng.uniform(0.5, 463, n))])
    exp = np.array([orc.geod_direct(*c)[:2] for c in x.T]).T
    for mode in (3, 4):
        got = _dev(mode, x)
        d = np.abs(got - exp)
        assert d[0].max() < 6e-15 and d[1].max() < 3e-14, (mode, d.max(1))   # Karney's own longitude noise near az 90 / 270
        worst = list(d.max(0).argsort()[-8:]) + list(range(8))
        for k in worst:                                                        # exact elliptic-integral quadrature
            ex = np.array(_exact_direct(*x[:, k]))
            assert np.abs(got[:, k] - ex).max() < 3e-15, (mode, x[:, k], got[:, k] - ex)


def test_inverse_exact_and_local_match_oracle():
    import oracle as orc
    rng = np.random.default_rng(2)
    for max_deg, tol_s, tol_a_m, loc_s, loc_a in ((0.06, 2e-8, 1e-7, 1e-5, 3e-8), (0.7, 2e-8, 1e-7, 2e-2, 1e-5)):
        x = _pairs(rng, 20000, max_deg)
        exp = np.array([orc.geod_inverse(*c)[:2] for c in x.T]).T
        ex = _dev(1, x)
        lo = _dev(2, x)
        far = exp[0] > 1.0
        dazi = lambda a: np.abs((a - exp[1] + 180) % 360 - 180)
        assert np.abs(ex[0] - exp[0]).max() < tol_s                        # metres
        assert (dazi(ex[1])[far] * np.pi / 180 * exp[0][far]).max() < tol_a_m  # cross-track metres
        # accuracy claim of geo::inverse_local (margins in hh_core.cuh are >= 100x these)
        assert np.abs(lo[0] - exp[0]).max() < loc_s
        assert dazi(lo[1])[far].max() < loc_a
    # degenerate: coincident points and meridian pairs
    x = np.array([[5.1, 7.1, 5.1, 7.1], [5.1, 7.1, 5.2, 7.1], [5.2, 7.1, 5.1, 7.1]]).T
    ex = _dev(1, x)
    assert ex[0, 0] == 0.0 and abs(ex[1, 1]) < 1e-12 and abs(abs(ex[1, 2]) - 180) < 1e-12
    assert abs(ex[0, 1] - orc.geod_inverse(5.1, 7.1, 5.2, 7.1)[0]) < 1e-8
