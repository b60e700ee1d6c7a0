"""The v4 step schedule (hhmarl_2d_b200/csrc/hh_v4.cuh) executed on the CPU by tests/emu against the reference's
golden trajectories and the C oracle -- the same bar as tests/test_gpu_parity.py (bit-exact bookkeeping, 1e-5
relative floats).  This pins the SEMANTICS of the kernel source (stage functions, draw order, hazards between roles)
without a GPU; the CUDA build itself is checked by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu_env  # noqa: E402
import golden_util as gu  # noqa: E402
import test_gpu_parity as tgp  # noqa: E402

pytestmark = pytest.mark.skipif(emu_env.cuda_include() is None, reason="needs cuda_runtime.h for the vector types")


@pytest.mark.parametrize("path", gu.golden_files(policy_levels=False), ids=lambda p: p.split("lowlevel_")[-1][:-4])
def test_v4_schedule_replays_reference_golden(path):
    with emu_env.emulated():
        tgp.test_cuda_replays_reference_golden(path)


def pursuit_actions(states, rng, p_random=0.15):
    """Actions that make the agents hunt: steer towards the nearest live opponent (15-degree heading steps), full speed
    when far, always fire cannon and missile.  Engagements -- cannon cones, rocket proximity, kills, friendly fire --
    then happen every few steps instead of once in a while as with uniformly random actions."""
    n = len(states)
    act = np.zeros((n, 2, 4), np.int32)
    for k, s in enumerate(states):
        for u in range(2):
            if rng.random() < p_random or not s.alive[u]:
                act[k, u] = (rng.integers(0, 13), rng.integers(0, 9), rng.integers(0, 2), rng.integers(0, 2))
                continue
            best, bd = None, 1e9
            for e in (2, 3):
                if s.alive[e]:
                    d = np.hypot(s.lon[e] - s.lon[u], s.lat[e] - s.lat[u])
                    if d < bd:
                        best, bd = e, d
            if best is None:
                act[k, u] = (6, 4, 1, 1)
                continue
            bearing = np.degrees(np.arctan2(s.lon[best] - s.lon[u], s.lat[best] - s.lat[u])) % 360
            err = (bearing - s.heading[u] + 180) % 360 - 180
            act[k, u] = (int(np.clip(round(err / 15) + 6, 0, 12)), 8 if bd > 0.03 else 3, 1, 1)
    act[:, 1, 3] = 0
    return act


def _many(level, mode, kw, n, T, reverse=False, arenas_per_cta=None, pursuit=False):
    import oracle as orc
    seed, base = 99173 + level, 1000
    with emu_env.emulated(reverse=reverse, arenas_per_cta=arenas_per_cta):
        env = tgp._vec(n, level, mode, seed, arena_base=base, autoreset=True, **kw)
        oracles = [orc.OracleEnv(orc.make_args(level=level, agent_mode=mode, **kw), seed, base + k) for k in range(n)]
        o1, o2 = env.reset_host()
        ref = [o.reset() for o in oracles]
        tgp._close(o1, np.stack([p[0] for p in ref]), "reset obs1")
        tgp._close(o2, np.stack([p[1] for p in ref]), "reset obs2")
        rng = np.random.default_rng(level)
        n_done = 0
        trace = []
        n_kill_steps = 0
        for t in range(T):
            if pursuit:
                act = pursuit_actions([o.state() for o in oracles], rng)
            else:
                act = np.stack([rng.integers(0, 13, (n, 2)), rng.integers(0, 9, (n, 2)), rng.integers(0, 2, (n, 2)),
                                rng.integers(0, 2, (n, 2))], axis=-1).astype(np.int32)
            alive_before = np.array([sum(o.state().alive[:4]) for o in oracles]) if pursuit else None
            g1, g2, grew, gdone = env.step_host(act)
            e1 = np.empty_like(g1); e2 = np.empty_like(g2); erew = np.empty((n, 2)); edone = np.empty(n, np.uint8)
            for k, o in enumerate(oracles):
                a1, a2, r, pres, d = o.step(act[k])
                edone[k], erew[k] = d, r
                if d:
                    a1, a2 = o.reset()
                e1[k], e2[k] = a1, a2
            assert (gdone == edone).all(), f"t={t} done mismatch at {np.nonzero(gdone != edone)[0][:8]}"
            tgp._close(grew, erew, f"t={t} rew")
            tgp._close(g1, e1, f"t={t} obs1")
            tgp._close(g2, e2, f"t={t} obs2")
            n_done += int(edone.sum())
            if pursuit:   # arena-steps in which an aircraft was lost (shot down or out of bounds); resets excluded
                after = np.array([sum(o.state().alive[:4]) for o in oracles])
                n_kill_steps += int(np.count_nonzero((after < alive_before) & (edone == 0)))
            trace.append((g1.copy(), g2.copy(), grew.copy(), gdone.copy()))
            if t % 50 == 49 or t == T - 1:
                st = env.get_state()
                os_ = [o.state() for o in oracles]
                for fld in tgp.F64_FIELDS:
                    tgp._close(st[fld], np.array([list(getattr(s, fld)[:4]) for s in os_]), f"t={t} {fld}")
                for fld in ("missile_remain", "missile_wait", "alive", "has_missile", "opp_to_attack", "cannon_remain",
                            "cannon_burst"):
                    assert (st[fld] == np.array([list(getattr(s, fld)[:4]) for s in os_])).all(), (t, fld)
                for fld in ("steps", "alive_agents", "alive_opps", "escaping", "escaping_time", "next_unit_id", "draws_g",
                            "draws_c"):
                    assert (st[fld] == np.array([getattr(s, fld) for s in os_])).all(), (t, fld)
                assert (st["error"] == 0).all()
        assert n_done > n // 2
        if pursuit:
            assert n_kill_steps > n, n_kill_steps           # losses in the middle of episodes, in every arena on average
        return trace


CASES = [(1, "fight", {}), (2, "fight", {}), (3, "fight", {}), (3, "escape", {"esc_dist_rew": True}),
         (3, "fight", {"glob_frac": 0.25, "friendly_punish": True, "rew_scale": 2}), (2, "fight", {"friendly_kill": False})]


@pytest.mark.parametrize("level,mode,kw", CASES)
def test_v4_schedule_matches_oracle(level, mode, kw):
    """77 arenas (ragged: 2 full CTAs + 13) x 330 ticks with in-step auto-reset against 77 scalar C oracles."""
    _many(level, mode, kw, n=77, T=330)


@pytest.mark.parametrize("level,mode,kw", [(3, "fight", {}), (2, "fight", {}), (3, "escape", {"esc_dist_rew": True}),
                                           (3, "fight", {"friendly_punish": True, "glob_frac": 0.5})])
def test_v4_schedule_matches_oracle_under_pursuit(level, mode, kw):
    """Same comparison with agents that hunt the opponents: dense cannon / rocket engagements and kills."""
    _many(level, mode, kw, n=48, T=400, pursuit=True)


def test_v4_schedule_has_no_order_dependence_between_threads_of_a_stage():
    """Roles of one stage run concurrently on the GPU: executing every role's threads in reverse order, and with a
    different number of arenas per CTA, must give bit-identical results."""
    a = _many(3, "fight", {}, n=40, T=200)
    b = _many(3, "fight", {}, n=40, T=200, reverse=True)
    c = _many(3, "fight", {}, n=40, T=200, arenas_per_cta=8)
    for x, y, z in zip(a, b, c):
        for p, q, r in zip(x, y, z):
            assert np.array_equal(p, q) and np.array_equal(p, r)


def test_direct_short_agrees_with_karney_direct():
    """geo::direct_short (one simulator tick: s12 <= 3 km near the map) against the full order-6 solver of the same
    header and against the oracle's restatement of geographiclib."""
    import ctypes
    import oracle as orc
    rng = np.random.default_rng(0)
    n = 20000
    q = np.empty((n, 4))
    q[:, 0] = rng.uniform(4.9, 5.7, n)
    q[:, 1] = rng.uniform(6.9, 7.7, n)
    q[:, 2] = np.where(rng.random(n) < 0.5, rng.integers(0, 360, n).astype(float), rng.uniform(-20, 400, n))
    q[:, 3] = np.where(rng.random(n) < 0.2, rng.uniform(250, 1030, n), rng.uniform(0.5, 463, n))
    out = {}
    with emu_env.emulated() as L:
        for mode in (0, 3, 4):
            o = np.empty((n, 2))
            assert L.hh_debug_geodesic(mode, n, q.ctypes.data, o.ctypes.data) == 0
            out[mode] = o
    for mode in (3, 4):
        d = np.abs(out[mode] - out[0])
        # direct() itself loses ~20 ulp of the longitude near azimuth 90 / 270 deg (difference of two O(0.1)
        # products in omg12); direct_short does not -- see the comparison with the exact solution below
        assert d[:, 0].max() < 6e-15 and d[:, 1].max() < 3e-14, (mode, d.max(0))
    ref = np.array([orc.geod_direct(*row)[:2] for row in q[:2000]])
    assert np.abs(out[3][:2000] - ref).max() < 3e-14
    # exact elliptic-integral quadrature (mpmath, 30 digits) at the 12 samples where the two solvers differ most
    # and 12 arbitrary ones
    from test_oracle_geodesic import _exact_direct
    worst = list(np.abs(out[3] - out[0]).max(1).argsort()[-12:]) + list(range(12))
    for k in worst:
        ex = np.array(_exact_direct(*q[k]))
        assert np.abs(out[3][k] - ex).max() < 2e-15, (q[k], out[3][k] - ex)
        assert np.abs(out[4][k] - ex).max() < 2e-15, (q[k], out[4][k] - ex)


# ------------------------------------------------------------------------------------------ scalar helpers (hh_core.cuh)
def _scalar(op, rows):
    import ctypes
    a = np.zeros((len(rows), 6))
    a[:, :np.asarray(rows).shape[1]] = rows
    out = np.empty((len(rows), 2))
    with emu_env.emulated() as L:
        L.hh_emu_scalar.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
        assert L.hh_emu_scalar(op, len(rows), a.ctypes.data, out.ctypes.data) == 0
    return out


def _ref_signed_heading_diff(actual, desired):        # warsim/utils/angles.py:22-29
    delta = desired - actual
    if delta < -180:
        delta = 360 + delta
    if delta > 180:
        delta = -360 + delta
    return delta


def _ref_normalize(a):                                 # warsim/utils/angles.py:10-15
    while a >= 360:
        a -= 360
    while a < 0:
        a += 360
    return a


def _ref_focus(h, lat_a, lon_a, lat_b, lon_b):         # envs/env_base.py:424-432 (numpy, as the reference computes it)
    th = ((90 - h) % 360) * (np.pi / 180)
    u = np.array([np.cos(th), np.sin(th)])
    v = np.array([lon_b - lon_a, lat_b - lat_a])
    x = np.clip(np.dot(u, v) / (np.linalg.norm(u) * np.linalg.norm(v) + 1e-10), -1, 1)
    return np.arccos(x) * 180 / np.pi


def test_scalar_helpers_follow_python_semantics():
    """The Python-semantics helpers of hh_core.cuh (the source the kernels compile) against CPython / numpy
    restatements of the reference formulas: float %, heading differences, radar gate, angle features."""
    import math
    rng = np.random.default_rng(11)
    # float % with positive modulus, incl. values next to multiples and the tiny-negative quirk
    xs = np.concatenate([rng.uniform(-1500, 1500, 20000), np.arange(-4, 5) * 360.0, np.arange(-4, 5) * 359.0,
                         np.nextafter(np.arange(1, 5) * 360.0, 0), np.nextafter(np.arange(1, 5) * 360.0, 1e9),
                         [-1e-20, -1e-300, 1e-20, -0.0, 0.0, 719.9999999999999, -359.99999999999994]])
    for m in (360.0, 359.0):
        got = _scalar(0, np.stack([xs, np.full_like(xs, m)], 1))[:, 0]
        want = np.array([x % m for x in xs.tolist()])
        assert np.array_equal(got, want), (m, xs[got != want][:5])
        gf = _scalar(11, np.stack([xs, np.full_like(xs, m)], 1))[:, 0]
        assert np.array_equal(gf, np.array([math.fmod(x, m) for x in xs.tolist()]))
    assert _scalar(0, [[-1e-20, 360.0]])[0, 0] == 360.0          # the reference's % returns the modulus itself here
    # signed_heading_diff / normalize_angle
    pairs = np.stack([rng.uniform(-50, 400, 5000), rng.uniform(-50, 400, 5000)], 1)
    pairs[:50] = np.round(pairs[:50])
    got = _scalar(1, pairs)[:, 0]
    assert np.array_equal(got, np.array([_ref_signed_heading_diff(a, b) for a, b in pairs.tolist()]))
    ang = np.concatenate([rng.uniform(-1000, 1000, 5000), [0.0, 360.0, -360.0, 359.99999999999994, 720.0]])
    assert np.array_equal(_scalar(2, ang[:, None])[:, 0], np.array([_ref_normalize(a) for a in ang.tolist()]))
    # radar gate (ac1.py:144-146): accepts relative bearings in (-1, 121) degrees
    hd = rng.uniform(0, 360, 20000); rel = rng.uniform(-180, 180, 20000)
    got = _scalar(3, np.stack([hd, (hd + rel) % 360], 1))[:, 0].astype(bool)
    clear = (np.abs(rel + 1) > 1e-9) & (np.abs(rel - 121) > 1e-9)
    assert np.array_equal(got[clear], ((rel > -1) & (rel < 121))[clear])
    # heading feature (sic: % 359 / 359), focus angle, heading difference, map helpers
    h = np.concatenate([rng.uniform(0, 360, 5000), np.arange(0, 361, 1.0)])
    want = np.clip((h % 359) / 359, 0, 1)
    assert np.abs(_scalar(4, h[:, None])[:, 0] - want).max() < 1e-15
    rows = np.stack([rng.uniform(0, 360, 5000), rng.uniform(5, 5.3, 5000), rng.uniform(7, 7.3, 5000),
                     rng.uniform(5, 5.3, 5000), rng.uniform(7, 7.3, 5000)], 1)
    rows[:10, 3:5] = rows[:10, 1:3] + 1e-6 * rng.standard_normal((10, 2))        # ~100 m apart: the + 1e-10 term matters
    got = _scalar(5, rows)[:, 0]
    want = np.array([_ref_focus(*r) for r in rows.tolist()])
    near = np.minimum(want, 180 - want) < 0.5      # arccos is ill-conditioned next to 0 / 180 deg: compare there loosely
    assert np.abs(got - want)[~near].max() < 1e-9 and np.abs(got - want)[near].max() < 2e-6
    sg = _scalar(6, rows)
    assert np.array_equal(sg[:, 0], sg[:, 1]) and set(np.unique(sg[:, 0])) <= {-1.0, 1.0}   # split form == one-piece form
    rp = _scalar(8, np.stack([rng.uniform(4.9, 5.4, 2000), rng.uniform(6.9, 7.4, 2000)], 1))
    assert rp.min() >= 0.0 and rp.max() <= 1.0
    assert [_scalar(7, [[k]])[0, 0] for k in (0, 5, 10, 11)] == [500.0, float.fromhex("0x1.8405555555554p+10"), 2000.0, 2000.0]
    if os.path.isdir("/root/reference"):            # this container only: the reference's own helpers
        sys.path.insert(0, "/root/reference/warsim")
        try:
            from utils.angles import signed_heading_diff, normalize_angle
            assert all(signed_heading_diff(a, b) == g for (a, b), g in zip(pairs[:500].tolist(), _scalar(1, pairs[:500])[:, 0]))
            assert all(normalize_angle(a) == g for a, g in zip(ang[:500].tolist(), _scalar(2, ang[:500, None])[:, 0]))
        finally:
            sys.path.pop(0)
