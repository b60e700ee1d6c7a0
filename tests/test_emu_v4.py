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


@pytest.mark.parametrize("path", gu.golden_files(), ids=lambda p: p.split("lowlevel_")[-1][:-4])
def test_v4_schedule_replays_reference_golden(path):
    with emu_env.emulated():
        tgp.test_cuda_replays_reference_golden(path)


def _many(level, mode, kw, n, T, reverse=False, arenas_per_cta=None):
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
        for t in range(T):
            act = np.stack([rng.integers(0, 13, (n, 2)), rng.integers(0, 9, (n, 2)), rng.integers(0, 2, (n, 2)),
                            rng.integers(0, 2, (n, 2))], axis=-1).astype(np.int32)
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
        return trace


CASES = [(1, "fight", {}), (2, "fight", {}), (3, "fight", {}), (3, "escape", {"esc_dist_rew": True}),
         (3, "fight", {"glob_frac": 0.25, "friendly_punish": True, "rew_scale": 2}), (2, "fight", {"friendly_kill": False})]


@pytest.mark.parametrize("level,mode,kw", CASES)
def test_v4_schedule_matches_oracle(level, mode, kw):
    """77 arenas (ragged: 2 full CTAs + 13) x 330 ticks with in-step auto-reset against 77 scalar C oracles."""
    _many(level, mode, kw, n=77, T=330)


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
