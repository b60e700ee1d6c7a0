"""hhmarl_2d_b200.evaluation on the CPU: the sums / report of evaluation.py:59-82 and the commander query loop of
evaluation.py:36-47 (batched over arenas) against a literal per-arena restatement."""
import json

import numpy as np
import torch

from hhmarl_2d_b200 import models as M
from hhmarl_2d_b200.env_hier import EVAL_INFO_KEYS
from hhmarl_2d_b200.evaluation import EvalStats, commander_actions


def test_eval_stats_report(tmp_path):
    rng = np.random.default_rng(0)
    st = EvalStats("cpu")
    ev = {k: 0 for k in EVAL_INFO_KEYS}
    ev["total_n_actions"] = 0
    n_evals = 0
    for _ in range(7):
        info = rng.integers(0, 4, (50, 12)).astype(np.int32)
        done = rng.integers(0, 2, 50).astype(np.uint8)
        count = rng.integers(0, 2, 50).astype(bool)
        st.update(torch.from_numpy(info), torch.from_numpy(done), torch.from_numpy(count))
        for a in range(50):          # evaluation.py:59-60, one arena-step at a time
            if count[a]:
                for k, v in zip(EVAL_INFO_KEYS, info[a]):
                    ev[k] += int(v)
                ev["total_n_actions"] += 1
                n_evals += int(done[a])
    tot = st.totals()
    assert tot["episodes"] == n_evals and all(tot[k] == ev[k] for k in ev)
    # postprocess_eval, evaluation.py:66-82
    want = {"win": ev["agents_win"] / n_evals * 100, "lose": ev["opps_win"] / n_evals * 100, "draw": ev["draw"] / n_evals * 100,
            "fight": ev["agent_fight"] / ev["agent_steps"] * 100, "esc": ev["agent_escape"] / ev["agent_steps"] * 100,
            "fight_opp": ev["opp_fight"] / ev["opp_steps"] * 100, "esc_opp": ev["opp_escape"] / ev["opp_steps"] * 100,
            "opp1": ev["opp1"] / ev["agent_fight"] * 100, "opp2": ev["opp2"] / ev["agent_fight"] * 100,
            "opp3": ev["opp3"] / ev["agent_fight"] * 100}
    path = tmp_path / "m.json"
    got = st.save(str(path))
    assert list(got) == list(want) and all(abs(got[k] - want[k]) < 1e-9 for k in want)
    assert json.loads(path.read_text()) == got
    # without `count` every arena-step enters the sums; empty denominators report 0 instead of raising
    st2 = EvalStats("cpu")
    assert st2.summary()["win"] == 0.0
    st2.update(torch.ones((4, 12), dtype=torch.int32), torch.tensor([1, 0, 0, 1], dtype=torch.uint8))
    assert st2.totals()["episodes"] == 2 and st2.totals()["draw"] == 4 and st2.totals()["total_n_actions"] == 4


def test_commander_actions_match_per_arena_loop():
    model = M.CommanderGru()
    M.fill_from_seed(model, 5, scale=1.0)   # a commander that uses all three actions
    model.eval()
    rng = np.random.default_rng(3)
    obs = torch.from_numpy(rng.random((9, 3, 34)).astype(np.float32))
    obs[2, 1] = 0                                    # a dead agent's observation is all zeros, it is still queried
    got = commander_actions(model, obs)
    assert got.dtype == torch.int32 and got.shape == (9, 3)
    for a in range(9):                               # evaluation.py:36-47 literally, batch 1
        states = [torch.zeros(1, 200), torch.zeros(1, 200)]
        for ag in range(3):
            d = {"obs_1_own": obs[a, ag][None], "obs_2": torch.zeros(1, 34), "obs_3": torch.zeros(1, 34),
                 "act_1_own": torch.zeros(1, 1), "act_2": torch.zeros(1, 1), "act_3": torch.zeros(1, 1)}
            with torch.no_grad():
                logits, states = model({"obs": d}, states, torch.ones(1, dtype=torch.int32))
            assert int(torch.argmax(logits[0])) == int(got[a, ag]), (a, ag)
    assert len(set(got.flatten().tolist())) > 1


class _FakeEnv:
    """Stands in for VecHighLevelEnv (same attributes evaluate() touches) with scripted episode lengths."""

    def __init__(self, n, lengths):
        self.n_arenas, self.dev, self.eval_info = n, torch.device("cpu"), True
        self.lengths = lengths            # lengths[a] = commander steps per episode of arena a
        self.t = torch.zeros(n, dtype=torch.int64)
        self.info = torch.zeros((n, 12), dtype=torch.int32)
        self.seen = []

    def reset(self):
        return torch.zeros((self.n_arenas, 3, 34))

    def step(self, act):
        self.seen.append(act.clone())
        self.t += 1
        done = (self.t % self.lengths == 0).to(torch.uint8)
        self.info.zero_()
        self.info[:, 2] = done.to(torch.int32)          # every episode ends as a draw
        self.info[:, 3] = (act > 0).sum(1)               # agent_fight
        self.info[:, 4] = (act == 0).sum(1)              # agent_escape
        self.info[:, 7] = 3
        return torch.zeros((self.n_arenas, 3, 34)), torch.zeros((self.n_arenas, 3)), done


def test_evaluate_counts_each_arenas_share_only():
    from hhmarl_2d_b200.evaluation import evaluate
    lengths = torch.tensor([2, 3, 5, 7])
    env = _FakeEnv(4, lengths)
    st = evaluate(env, None, n_episodes=10)              # share = ceil(10 / 4) = 3 episodes per arena
    ev = st.totals()
    assert ev["episodes"] == 12 and ev["draw"] == 12
    assert ev["total_n_actions"] == int((3 * lengths).sum())      # steps past an arena's share are not counted
    assert ev["agent_steps"] == 3 * ev["total_n_actions"] and ev["agent_escape"] == 0
    assert len(env.seen) == 21 and all((a == 1).all() for a in env.seen)   # no commander: closest opponent
    assert st.summary()["draw"] == 100.0 and st.summary()["fight"] == 100.0
    env.eval_info = False
    try:
        evaluate(env, None, 4)
        raise AssertionError("evaluate() must refuse an env without eval_info")
    except ValueError:
        pass
    # with a commander the actions come from the model, one row per arena
    model = M.CommanderGru()
    M.fill_from_seed(model, 5, scale=1.0)
    env2 = _FakeEnv(4, lengths)
    st2 = evaluate(env2, model.eval(), n_episodes=4)
    assert st2.totals()["episodes"] == 4 and env2.seen[0].shape == (4, 3) and env2.seen[0].dtype == torch.int32


def test_policy_key_selects_opponent_fight_policies_only_in_lowlevel_evaluation():
    """env_base.py:385-390."""
    from argparse import Namespace
    from hhmarl_2d_b200.env_hier import VecHighLevelEnv, make_hier_args
    env = VecHighLevelEnv.__new__(VecHighLevelEnv)       # no device needed for the key logic
    env.args = make_hier_args()
    assert [env._policy_key(m, ac, f) for m in ("fight", "escape") for ac in (1, 2) for f in (0, 3)] == \
           ["fight_1", "fight_1", "fight_2", "fight_2", "escape_1", "escape_1", "escape_2", "escape_2"]
    env.args = Namespace(**{**vars(make_hier_args()), "eval_hl": False})
    assert [env._policy_key(m, ac, f) for m in ("fight", "escape") for ac in (1, 2) for f in (0, 3)] == \
           ["fight_1", "fight_1_opp", "fight_2", "fight_2_opp", "escape_1", "escape_1", "escape_2", "escape_2"]
    env._h = None
