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
