"""The C oracle (oracle/hhmarl_oracle.c) replayed against the committed golden trajectories of the
reference's unmodified LowLevelEnv (tests/golden/gen_golden.py).  Discrete bookkeeping must be
identical; floats agree to 1e-9 (numpy's BLAS dot / norm differ from scalar C in the last ulp,
which acos amplifies near 0 and 180 degrees)."""
import numpy as np
import pytest

import golden_util as gu
import oracle as orc


@pytest.mark.parametrize("path", gu.golden_files(), ids=lambda p: p.split("lowlevel_")[-1][:-4])
def test_oracle_replays_golden(path):
    g, seed, arena, level, mode, kw = gu.load(path)
    pol = gu.GoldenPolicy(g) if level >= 4 else None     # levels 4/5: queries checked, recorded answers fed back
    env = orc.OracleEnv(orc.make_args(level=level, agent_mode=mode, **kw), seed, arena, policy_fn=pol.fn if pol else None)
    o1, o2 = env.reset()
    ep = 0
    np.testing.assert_allclose(o1, g["reset_obs1"][0], atol=1e-7)
    np.testing.assert_allclose(o2, g["reset_obs2"][0], atol=1e-7)
    n_kill_steps = 0
    for t in range(len(g["done"])):
        if pol:
            pol.begin_step(t)
        o1, o2, rew, present, done = env.step(g["actions"][t])
        if pol:
            pol.end_step()
        st = env.state()
        assert st.error == 0
        assert done == bool(g["done"][t]), t
        assert (present == g["present"][t]).all(), t
        sc = [st.steps, st.alive_agents, st.alive_opps, st.escaping, st.escaping_time, st.next_unit_id,
              st.draws_g, st.draws_c]
        assert sc == list(g["scalars"][t]), (t, sc, g["scalars"][t])
        for k in gu.I32 + gu.RI32:
            assert (np.array(getattr(st, k)[:4]) == g[k][t]).all(), (t, k)
        for k in gu.F64 + gu.R64:
            np.testing.assert_allclose(np.array(getattr(st, k)[:4]), g[k][t], rtol=1e-12, atol=1e-9, err_msg=f"{t} {k}")
        np.testing.assert_allclose(o1, g["obs1"][t], atol=1e-7)
        np.testing.assert_allclose(o2, g["obs2"][t], atol=1e-7)
        np.testing.assert_allclose(rew, g["rew"][t], rtol=1e-9, atol=1e-9)
        if t and (g["alive"][t] != g["alive"][t - 1]).any():
            n_kill_steps += 1
        if done:
            ep += 1
            o1, o2 = env.reset()
            np.testing.assert_allclose(o1, g["reset_obs1"][ep], atol=1e-7)
            np.testing.assert_allclose(o2, g["reset_obs2"][ep], atol=1e-7)
    assert ep == len(g["reset_obs1"]) - 1 and n_kill_steps > 0
    if level == 5 and mode == "fight":
        assert set(g["c_pset"][g["c_unit"] > 0].tolist()) == {3, 4, 5}      # L3-fight, L4-fight and L3-escape opponent sets
