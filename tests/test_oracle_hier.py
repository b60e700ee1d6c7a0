"""C oracle of the reference's HighLevelEnv (oracle/hhmarl_oracle.c: orc_hier_*) against the golden
trajectories of the unmodified envs/env_hier.py (tests/golden/gen_golden_hier.py) and, where the reference is
present, live."""
import ast
import glob
import os

import numpy as np
import pytest

import oracle as orc
import ref_harness as rh

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "hier_*.npz")))


@pytest.mark.parametrize("path", FILES, ids=lambda p: os.path.basename(p)[:-4])
def test_hier_oracle_replays_golden(path):
    g = dict(np.load(path))
    seed, arena = (int(x) for x in g["meta"])
    kw = ast.literal_eval(str(g["kw"]))
    cur = {"t": 0, "k": 0}

    def pol(unit_id, ac_type, mode, pset, obs):
        t, k = cur["t"], cur["k"]
        assert k < g["n_calls"][t], "more policy queries than the reference made"
        assert (unit_id, ac_type, mode) == (g["c_unit"][t][k], g["c_type"][t][k], g["c_mode"][t][k]), (t, k)
        np.testing.assert_allclose(obs, g["c_obs"][t][k][:len(obs)], atol=1e-6)
        cur["k"] += 1
        return g["c_act"][t][k][:4 if ac_type == 1 else 3]

    env = orc.OracleHierEnv(orc.make_hier_args(**kw), seed, arena, pol)
    np.testing.assert_allclose(env.reset(), g["resets"][0], atol=1e-7)
    ep = 0
    for t in range(len(g["done"])):
        cur["t"], cur["k"] = t, 0
        obs, rew, done, info = env.step(g["ca"][t])
        st = env.state()
        assert st.error == 0 and cur["k"] == g["n_calls"][t]
        assert done == bool(g["done"][t]) and info[0] == g["subs"][t]
        assert list(info[1:7]) == list(g["ca_out"][t])
        assert [st.steps, st.alive_agents, st.alive_opps, st.next_unit_id, st.draws_g, st.draws_c] == list(g["scalars"][t])
        np.testing.assert_allclose(obs, g["obs"][t], atol=1e-6)
        np.testing.assert_allclose(rew, g["rew"][t], atol=1e-9)
        if "eval" in g:   # args.eval_info: the info dict of env_base.py:91-107, key by key
            assert list(env.eval_info().values()) == list(g["eval"][t]), t
        if done:
            ep += 1
            np.testing.assert_allclose(env.reset(), g["resets"][ep], atol=1e-7)
    assert ep == len(g["resets"]) - 1 and ep >= 3
    if "eval" in g:   # the fixture holds wins, losses and draws
        assert tuple(env.eval_info()) == orc.EVAL_INFO_KEYS and all(g["eval"][:, k].sum() > 0 for k in (0, 1, 2))


@pytest.mark.skipif(not rh.reference_available(), reason="/root/reference not present")
def test_hier_oracle_matches_reference_live():
    def policy(u, t, m, o):
        h = int(np.abs(np.round(np.asarray(o, np.float64), 4)).sum() * 1e4) + 3 * u + m
        heads = (13, 9, 2, 2) if t == 1 else (13, 9, 2)
        return [(h // (1 + 2 * k)) % n for k, n in enumerate(heads)]

    ref = rh.ReferenceHierEnv(rh.make_hier_namespace(), 77, 5, policy)
    oe = orc.OracleHierEnv(orc.make_hier_args(), 77, 5, lambda u, t, m, ps, o: policy(u, t, m, o))
    rng = np.random.default_rng(0)
    np.testing.assert_allclose(ref.reset(), oe.reset(), atol=1e-7)
    for t in range(60):
        ca = rng.integers(0, 3, 3)
        o, r, d, ns, ca_out = ref.step(ca)
        o2, r2, d2, info = oe.step(ca)
        assert d == d2 and ns == info[0]
        np.testing.assert_allclose(o, o2, atol=1e-6)
        np.testing.assert_allclose(r, r2, atol=1e-9)
        if d:
            np.testing.assert_allclose(ref.reset(), oe.reset(), atol=1e-7)
