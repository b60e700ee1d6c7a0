"""Pins the oracle's Karney restatement (oracle/geodesic.c) -- the reference has no tests and
geographiclib is absent, so the pins are geographiclib's documented known answers, the session
spot values of SURVEY.md Appendix B, and 30-digit mpmath quadrature of the exact integrals."""
import numpy as np
import pytest

import oracle as orc


def test_known_answers_geographiclib_docs():
    lat2, lon2, azi2 = orc.geod_direct(40.6, -73.8, 51, 5.5e6)
    assert abs(lat2 - 51.884564505606) < 1e-11
    assert abs(lon2 - (-1.141172861200)) < 1e-11
    assert abs(azi2 - 107.189397162606) < 1e-11
    s12, azi1, azi2 = orc.geod_inverse(-41.32, 174.81, 40.96, -5.50)
    assert abs(s12 - 19959679.267) < 1e-3
    assert abs(azi1 - 161.067669986160) < 1e-10
    assert abs(azi2 - 18.825195123248) < 1e-10


def test_local_regime_spot_values():
    s12, azi1, _ = orc.geod_inverse(5.1, 7.1, 5.12, 7.13)
    assert abs(s12 - 3994.5443109607) < 1e-8
    assert abs(azi1 - 56.379467931477) < 1e-10
    lat2, lon2, _ = orc.geod_direct(5.1, 7.1, 33, 463)
    assert abs(lat2 - 5.103511424264236) < 1e-14
    assert abs(lon2 - 7.102274218421746) < 1e-14


def _exact_direct(lat1, lon1, azi1, s12):
    """Exact (quadrature) solution of the direct problem on WGS84 with mpmath."""
    import mpmath as mp
    mp.mp.dps = 30
    a = mp.mpf(6378137)
    f = 1 / mp.mpf("298.257223563")
    b = a * (1 - f)
    ep2 = f * (2 - f) / (1 - f) ** 2
    phi1, al1 = mp.radians(mp.mpf(lat1)), mp.radians(mp.mpf(azi1))
    beta1 = mp.atan((1 - f) * mp.tan(phi1))
    sa0 = mp.sin(al1) * mp.cos(beta1)
    ca0 = mp.sqrt(1 - sa0 ** 2)
    sig1 = mp.atan2(mp.sin(beta1), mp.cos(al1) * mp.cos(beta1))
    om1 = mp.atan2(sa0 * mp.sin(sig1), mp.cos(sig1))
    k2 = ep2 * ca0 ** 2
    dist = lambda s2: b * mp.quad(lambda s: mp.sqrt(1 + k2 * mp.sin(s) ** 2), [sig1, s2])
    sig2 = mp.findroot(lambda s2: dist(s2) - s12, sig1 + mp.mpf(s12) / b)
    om2 = mp.atan2(sa0 * mp.sin(sig2), mp.cos(sig2))
    I3 = mp.quad(lambda s: (2 - f) / (1 + (1 - f) * mp.sqrt(1 + k2 * mp.sin(s) ** 2)), [sig1, sig2])
    lam12 = (om2 - om1) - f * sa0 * I3
    beta2 = mp.asin(ca0 * mp.sin(sig2))
    phi2 = mp.atan(mp.tan(beta2) / (1 - f))
    return float(mp.degrees(phi2)), float(lon1 + mp.degrees(lam12))


@pytest.mark.parametrize("case", [
    (5.1, 7.1, 33.0, 463.0), (5.25, 7.02, 271.5, 51.4444), (5.0, 7.3, 0.0, 1028.888),
    (5.17, 7.21, 180.0, 200.0), (5.29, 7.0, 123.456, 25000.0), (5.01, 7.29, 359.0, 49999.0),
])
def test_direct_and_inverse_vs_quadrature(case):
    lat1, lon1, azi1, s12 = case
    lat2, lon2, _ = orc.geod_direct(lat1, lon1, azi1, s12)
    elat2, elon2 = _exact_direct(lat1, lon1, azi1, s12)
    assert abs(lat2 - elat2) < 5e-14 and abs(lon2 - elon2) < 5e-14
    s, a1, _ = orc.geod_inverse(lat1, lon1, elat2, elon2)
    assert abs(s - s12) <= 2e-9 * max(1.0, s12 / 1000)      # nanometres
    da = (a1 - azi1 + 180) % 360 - 180
    assert abs(da) * np.pi / 180 * s12 < 1e-7                # < 0.1 micron of cross-track


def test_inverse_degenerate_cases():
    s, a1, _ = orc.geod_inverse(5.1, 7.1, 5.1, 7.1)        # coincident points
    assert s == 0.0
    s, a1, _ = orc.geod_inverse(5.1, 7.1, 5.2, 7.1)        # meridian
    assert abs(a1) < 1e-12 and abs(s - 11057.4) < 5
    s2, a2, _ = orc.geod_inverse(5.2, 7.1, 5.1, 7.1)
    assert abs(abs(a2) - 180) < 1e-12 and abs(s - s2) < 1e-9


def test_philox_known_answers_and_mapping():
    # Random123 kat_vectors: philox4x32-10
    assert orc.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert orc.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (
        0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert orc.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == (
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    s = orc.PhiloxStream(7, 3, 0)
    xs = [s.random() for _ in range(1000)]
    assert 0.0 <= min(xs) and max(xs) < 1.0 and 0.45 < np.mean(xs) < 0.55
    s = orc.PhiloxStream(7, 3, 0)
    assert [s.randint(1, 2) for _ in range(50)].count(1) > 10
