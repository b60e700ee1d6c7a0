"""Evaluation statistics of the commander environment (the reference's evaluation.py) for N arenas at once.

The reference plays N_EVALS = 1000 episodes one after the other; every step it adds the env's `info` dict
(args.eval_info, env_base.py:91-107) to running sums (evaluation.py:59-60) and finally reports win / lose / draw
and fight / escape percentages as JSON (postprocess_eval, evaluation.py:66-82).  Here the arenas of a
VecHighLevelEnv play in lock-step with auto-reset, `VecHighLevelEnv.info` (csrc/hh_hier.cu: eval_info_kernel)
holds the same twelve counters per arena and step, and EvalStats sums them on the device.
"""
from __future__ import annotations

import json

from .env_hier import EVAL_INFO_KEYS, OBS_HL, VecHighLevelEnv


class EvalStats:
    """Running sums of evaluation.py:59-60 over arenas and steps; `summary()` is postprocess_eval
    (evaluation.py:66-82) with N_EVALS = the number of finished episodes."""

    def __init__(self, device):
        import torch
        self._torch = torch
        self.sums = torch.zeros(len(EVAL_INFO_KEYS), dtype=torch.int64, device=device)
        self.episodes = torch.zeros((), dtype=torch.int64, device=device)
        self.total_n_actions = 0

    def update(self, info, done, count=None):
        """info int32 [N,12], done u8 [N] of one commander step; `count` (bool [N]) restricts the step to the arenas
        whose episode still counts (evaluate() stops an arena after its share of episodes)."""
        if count is None:
            self.sums += info.sum(dim=0, dtype=self._torch.int64)
            self.episodes += (done != 0).sum()
            self.total_n_actions += int(info.shape[0])
        else:
            self.sums += (info * count[:, None]).sum(dim=0, dtype=self._torch.int64)
            self.episodes += ((done != 0) & count).sum()
            self.total_n_actions += int(count.sum())

    def totals(self) -> dict:
        ev = dict(zip(EVAL_INFO_KEYS, (int(v) for v in self.sums.cpu())))
        ev["total_n_actions"] = self.total_n_actions
        ev["episodes"] = int(self.episodes)
        return ev

    def summary(self) -> dict:
        ev = self.totals()

        def pct(a, b):   # the reference divides unguarded (evaluation.py:68-77); an empty denominator reports 0
            return (a / b) * 100 if b else 0.0

        return {"win": pct(ev["agents_win"], ev["episodes"]), "lose": pct(ev["opps_win"], ev["episodes"]),
                "draw": pct(ev["draw"], ev["episodes"]),
                "fight": pct(ev["agent_fight"], ev["agent_steps"]), "esc": pct(ev["agent_escape"], ev["agent_steps"]),
                "fight_opp": pct(ev["opp_fight"], ev["opp_steps"]), "esc_opp": pct(ev["opp_escape"], ev["opp_steps"]),
                "opp1": pct(ev["opp1"], ev["agent_fight"]), "opp2": pct(ev["opp2"], ev["agent_fight"]),
                "opp3": pct(ev["opp3"], ev["agent_fight"])}

    def save(self, path: str) -> dict:
        """Metrics_<config>.json of the reference (evaluation.py:80-81: json.dump(evals, file, indent=3))."""
        evals = self.summary()
        with open(path, "w") as f:
            json.dump(evals, f, indent=3)
        return evals


def commander_actions(model, obs):
    """evaluation.py:36-47 for every arena: the agents are queried in id order, each with ONLY its own observation
    (cc_obs, evaluation.py:20-28: team-mates' observations and all actions zero), deterministic (explore=False ->
    argmax); the GRU state starts at zero EVERY step and is handed from one agent to the next (states[0] / states[1]
    are overwritten inside the loop over agents) -- reproduced as is.  obs f32 [N,3,34] -> int32 [N,3]."""
    import torch
    n = obs.shape[0]
    dev = obs.device
    h = [torch.zeros((n, 200), device=dev), torch.zeros((n, 200), device=dev)]
    z1, zo = torch.zeros((n, 1), device=dev), torch.zeros((n, OBS_HL), device=dev)
    ones = torch.ones(n, dtype=torch.int32)
    out = torch.empty((n, 3), dtype=torch.int32, device=dev)
    with torch.no_grad():
        for ag in range(3):
            d = {"obs_1_own": obs[:, ag], "obs_2": zo, "obs_3": zo, "act_1_own": z1, "act_2": z1, "act_3": z1}
            logits, h = model({"obs": d}, h, ones)
            out[:, ag] = torch.argmax(logits, dim=-1).to(torch.int32)
    return out


def evaluate(env: VecHighLevelEnv, model=None, n_episodes: int = 1000, max_steps: int = 100000) -> EvalStats:
    """evaluation.py:30-63, 104-110 on all arenas at once.  `model` = a CommanderGru (args.eval_hl) or None: every
    agent attacks its closest opponent (evaluation.py:48-51).  Arena a plays episodes until it has finished its share
    ceil(n_episodes / N); steps of arenas that are past their share are not counted, so exactly
    N * ceil(n_episodes / N) >= n_episodes episodes enter the sums."""
    import torch
    if not env.eval_info:
        raise ValueError("evaluate(): the env must be created with args.eval_info = True (config.py:49)")
    n, dev = env.n_arenas, env.dev
    share = -(-int(n_episodes) // n)
    played = torch.zeros(n, dtype=torch.int64, device=dev)
    stats = EvalStats(dev)
    obs = env.reset()
    fixed = torch.ones((n, 3), dtype=torch.int32, device=dev)
    for _ in range(max_steps):
        count = played < share
        if not bool(count.any()):
            break
        act = commander_actions(model, obs) if model is not None else fixed
        obs, _, done = env.step(act)
        stats.update(env.info, done, count)
        played += ((done != 0) & count).to(torch.int64)
    return stats
