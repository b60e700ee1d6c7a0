"""Live pin of the C oracle against the reference's own files (imported unmodified from
/root/reference under stubs, oracle/ref_harness.py).  Skipped where the reference is absent
(the GPU box); the committed golden files carry the same evidence there."""
import numpy as np
import pytest

import oracle as orc
import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="/root/reference not present")


@pytest.mark.parametrize("level,mode,arena,kw", [
    (1, "fight", 11, {}), (2, "fight", 12, {}), (3, "fight", 13, {}), (3, "fight", 14, {"glob_frac": 0.3}),
    (3, "escape", 15, {"esc_dist_rew": True}),
])
def test_oracle_matches_reference_live(level, mode, arena, kw):
    seed, n = 424242, 400
    ref = rh.ReferenceEnv(rh.make_namespace(level=level, agent_mode=mode, **kw), seed, arena)
    oe = orc.OracleEnv(orc.make_args(level=level, agent_mode=mode, **kw), seed, arena)
    rng = np.random.default_rng(arena)
    a, b = ref.reset(), oe.reset()
    np.testing.assert_allclose(a[0], b[0], atol=1e-7)
    np.testing.assert_allclose(a[1], b[1], atol=1e-7)
    for t in range(n):
        act = np.stack([rng.integers(0, [13, 9, 2, 2]), rng.integers(0, [13, 9, 2, 2])]).astype(np.int32)
        o1, o2, r, p, d = ref.step(act)
        q1, q2, r2, p2, d2 = oe.step(act)
        rs, st = ref.state(), oe.state()
        assert d == d2 and (p == p2).all()
        assert list(rs["scalars"]) == [st.steps, st.alive_agents, st.alive_opps, st.escaping, st.escaping_time,
                                       st.next_unit_id, st.draws_g, st.draws_c]
        for k in ("alive", "missile_remain", "missile_wait", "has_missile", "opp_to_attack", "r_alive", "r_age"):
            assert (rs[k] == np.array(getattr(st, k)[:4])).all(), (t, k)
        for k in ("lat", "lon", "heading", "speed", "new_heading", "cannon_remain", "r_lat", "r_lon", "r_heading"):
            np.testing.assert_allclose(rs[k], np.array(getattr(st, k)[:4]), rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(o1, q1, atol=1e-7)
        np.testing.assert_allclose(o2, q2, atol=1e-7)
        np.testing.assert_allclose(r, r2, atol=1e-9)
        if d:
            a, b = ref.reset(), oe.reset()
            np.testing.assert_allclose(a[0], b[0], atol=1e-7)


@pytest.mark.parametrize("level,mode,arena,kw", [(4, "fight", 21, {}), (5, "fight", 22, {}), (5, "fight", 23, {"glob_frac": 0.2}),
                                                 (5, "escape", 24, {"esc_dist_rew": True})])
def test_oracle_matches_reference_live_frozen_policy_levels(level, mode, arena, kw):
    """Levels 4/5 (env_base.py:349-398, env_hetero.py:49-59, ammunition env_base.py:566-578): the reference's frozen
    policies are stubbed at the call site by a deterministic function of the query; the oracle's callback must see the
    SAME queries (unit, type, mode, policy set k, mid-step observation) in the same order."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from gen_golden import pseudo_policy
    seed, n = 515151, (1500 if (level, mode) == (5, "fight") else 700)
    ref_calls, orc_calls = [], []

    def ref_fn(u, t, m, ps, o):
        ref_calls.append((u, t, m, ps, np.array(o, np.float32)))
        return pseudo_policy(u, t, m, ps, o)

    def orc_fn(u, t, m, ps, o):
        # answer from the REFERENCE's observation of the same query so that a 1e-7 float difference cannot fork the runs
        k = len(orc_calls)
        orc_calls.append((u, t, m, ps, np.array(o, np.float32)))
        assert k < len(ref_calls)
        return np.asarray(pseudo_policy(u, t, m, ps, ref_calls[k][4]), np.int32)

    ref = rh.ReferenceEnv(rh.make_namespace(level=level, agent_mode=mode, **kw), seed, arena, policy_fn=ref_fn)
    oe = orc.OracleEnv(orc.make_args(level=level, agent_mode=mode, **kw), seed, arena, policy_fn=orc_fn)
    rng = np.random.default_rng(arena)
    a, b = ref.reset(), oe.reset()
    np.testing.assert_allclose(a[0], b[0], atol=1e-7)
    psets, episodes, n_q = set(), 0, 0
    for t in range(n):
        act = np.stack([rng.integers(0, [13, 9, 2, 2]), rng.integers(0, [13, 9, 2, 2])]).astype(np.int32)
        ref_calls.clear(); orc_calls.clear()
        o1, o2, r, p, d = ref.step(act)
        q1, q2, r2, p2, d2 = oe.step(act)
        assert len(ref_calls) == len(orc_calls)
        for x, y in zip(ref_calls, orc_calls):
            assert x[:4] == y[:4], (t, x[:4], y[:4])
            np.testing.assert_allclose(x[4], y[4], atol=1e-6)
            psets.add(x[3]); n_q += 1
        rs, st = ref.state(), oe.state()
        assert d == d2 and (p == p2).all()
        assert list(rs["scalars"]) == [st.steps, st.alive_agents, st.alive_opps, st.escaping, st.escaping_time,
                                       st.next_unit_id, st.draws_g, st.draws_c]
        for k in ("alive", "missile_remain", "rocket_max", "missile_wait", "has_missile", "opp_to_attack", "r_alive", "r_age"):
            assert (rs[k] == np.array(getattr(st, k)[:4])).all(), (t, k)
        for k in ("lat", "lon", "heading", "speed", "new_heading", "cannon_remain", "cannon_max", "r_lat", "r_lon", "r_heading"):
            np.testing.assert_allclose(rs[k], np.array(getattr(st, k)[:4]), rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(o1, q1, atol=1e-7)
        np.testing.assert_allclose(o2, q2, atol=1e-7)
        np.testing.assert_allclose(r, r2, atol=1e-9)
        if d:
            episodes += 1
            a, b = ref.reset(), oe.reset()
            np.testing.assert_allclose(a[0], b[0], atol=1e-7)
            np.testing.assert_allclose(a[1], b[1], atol=1e-7)
    assert episodes >= 4 and n_q > 800
    if level == 5 and mode == "fight":
        assert psets == {3, 4, 5}
    else:
        assert psets == {0}
