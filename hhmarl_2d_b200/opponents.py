"""Frozen-policy opponents for levels 4 and 5 (fictitious self-play), batched across arenas.

Mirrors HHMARLBaseEnv._get_policies / _policy_actions (env_base.py:312-398): the opponent's own observation
(friend observation and all actions zero) goes through the frozen actor, and each MultiDiscrete head takes its
argmax.  The reference does this with one batch-1 forward per opponent per step; here every arena's opponent 3
(AC1) and opponent 4 (AC2) go through one batched forward per policy set.

Policy containers follow the reference's dict shapes:
  level 4           : {"fight_1": Fight1, "fight_2": Fight2}                                  (L3 fight policies)
  level 5, fight    : {3: {"fight_1", "fight_2"}, 4: {"fight_1", "fight_2"}, 5: {"escape_1", "escape_2"}}
  level 5, escape   : {"fight_1": Fight1, "fight_2": Fight2}                                  (L5 fight policies)
The reference torch.load()s whole-module pickles of RLlib models from policies/ (not in the repository and not
loadable without ray); here the containers hold the ray-free restatements (hhmarl_2d_b200.models), whose
state_dict keys equal the reference's, initialised from a seed when no weights are given.
"""
from __future__ import annotations

import ctypes

import torch

from . import _native as nat
from . import models as M

OBS_FIGHT = {1: M.OBS_AC1, 2: M.OBS_AC2}
OBS_ESC = {1: M.OBS_ESC_AC1, 2: M.OBS_ESC_AC2}


def default_policies(level: int, agent_mode: str, seed: int = 0, device="cpu"):
    """Seeded stand-ins with the reference's container shape (there are no trained weights in this repo)."""
    def pair(mode, s):
        a, b = M.build_policy_pair(mode)
        M.fill_from_seed(a, s)
        M.fill_from_seed(b, s + 1)
        return a.to(device).eval(), b.to(device).eval()

    if level == 5 and agent_mode == "fight":
        out = {}
        for k in (3, 4):
            f1, f2 = pair("fight", seed + 10 * k)
            out[k] = {"fight_1": f1, "fight_2": f2}
        e1, e2 = pair("escape", seed + 50)
        out[5] = {"escape_1": e1, "escape_2": e2}
        return out
    f1, f2 = pair("fight", seed + 30)
    return {"fight_1": f1, "fight_2": f2}


class OpponentPolicies:
    def __init__(self, level: int, agent_mode: str, policies=None, seed: int = 0, device="cuda"):
        self.level, self.agent_mode = level, agent_mode
        self.device = torch.device(device)
        self.policies = policies if policies is not None else default_policies(level, agent_mode, seed, self.device)
        self.per_set = level == 5 and agent_mode == "fight"
        # user-supplied policies (checkpoint.load_opponent_policies defaults to the CPU) move to the env's device: the
        # fused kernel is handed their data_ptr()s
        for d in (self.policies.values() if self.per_set else (self.policies,)):
            for m in d.values():
                m.to(self.device).eval()

    def _models_for(self, k):
        """(model for opponent id 3 / AC1, model for opponent id 4 / AC2, mode) for policy set k."""
        if self.per_set:
            d = self.policies[k]
            if k == 5:
                return d["escape_1"], d["escape_2"], "escape"
            return d["fight_1"], d["fight_2"], "fight"
        return self.policies["fight_1"], self.policies["fight_2"], "fight"

    # ---- fused path: every (policy set, opponent) actor is one chain of ONE hh_policy_forward_ex launch
    def _fused_setup(self, n):
        from .fused_forward import FusedActor
        sets = (3, 4, 5) if self.per_set else (0,)
        self._fa = {k: tuple(FusedActor(m) for m in self._models_for(k)[:2]) for k in sets}
        dev = self.device
        self._fa_act = torch.zeros((n, 2, 4), dtype=torch.int32, device=dev)
        self._fa_rows = torch.zeros((3, n), dtype=torch.int32, device=dev)
        self._fa_ranges = torch.zeros((3, 2), dtype=torch.int32, device=dev)
        self._fa_keys = (ctypes.c_int32 * 3)(3, 4, 5)
        self._fa_n = n

    @torch.no_grad()
    def act_fused(self, opp_obs3, opp_obs4, policy_set=None, precision: int = 2):
        """Same result as act() (per-head argmax of the frozen actors) in three launches and without host
        synchronisation: hh_policy_rows_by_key lists the arenas of every policy set on the device, then all
        (set, opponent) actors run as chains of ONE hh_policy_forward_ex launch that writes the actions in place."""
        from .fused_forward import run_chains
        n = opp_obs3.shape[0]
        if getattr(self, "_fa_n", None) != n:
            self._fused_setup(n)
        act = self._fa_act
        p3, p4 = act.data_ptr(), act.data_ptr() + 16           # [n, 2, 4] int32: opponent id 3 / id 4 of every arena
        fills = []
        if self.per_set:
            assert policy_set.dtype == torch.uint8 and policy_set.is_contiguous()
            rows, ranges = self._fa_rows, self._fa_ranges
            st = torch.cuda.current_stream(self.device).cuda_stream
            nat.check(nat.lib().hh_policy_rows_by_key(n, policy_set.data_ptr(), 3, self._fa_keys, rows.data_ptr(),
                                                      ranges.data_ptr(), st), "hh_policy_rows_by_key")
            for j, k in enumerate((3, 4, 5)):
                f1, f2 = self._fa[k]
                fills.append(lambda c, f=f1, j=j: f.fill_chain(c, opp_obs3, n, rows=rows, range_dev=ranges[j], act_ptr=p3, ld_act=2))
                fills.append(lambda c, f=f2, j=j: f.fill_chain(c, opp_obs4, n, rows=rows, range_dev=ranges[j], act_ptr=p4, ld_act=2))
        else:
            f1, f2 = self._fa[0]
            fills.append(lambda c: f1.fill_chain(c, opp_obs3, n, act_ptr=p3, ld_act=2))
            fills.append(lambda c: f2.fill_chain(c, opp_obs4, n, act_ptr=p4, ld_act=2))
        run_chains(fills, self.device, precision)
        return act

    @torch.no_grad()
    def act(self, opp_obs3: torch.Tensor, opp_obs4: torch.Tensor, policy_set: torch.Tensor | None = None,
            return_logits: bool = False):
        """opp_obs3 [N,30], opp_obs4 [N,29] (fight rows use the first 26 / 24) -> int32 [N,2,4]."""
        n = opp_obs3.shape[0]
        act = torch.zeros((n, 2, 4), dtype=torch.int32, device=opp_obs3.device)
        logits_out = {}
        sets = (3, 4, 5) if self.per_set else (0,)
        for k in sets:
            m1, m2, mode = self._models_for(k)
            d1, d2 = (OBS_FIGHT[1], OBS_FIGHT[2]) if mode == "fight" else (OBS_ESC[1], OBS_ESC[2])
            if self.per_set:
                idx = torch.nonzero(policy_set == k, as_tuple=False).flatten()
                if idx.numel() == 0:
                    continue
                o3, o4 = opp_obs3.index_select(0, idx)[:, :d1], opp_obs4.index_select(0, idx)[:, :d2]
            else:
                idx, o3, o4 = None, opp_obs3[:, :d1], opp_obs4[:, :d2]
            l3, l4 = m1.actor(o3), m2.actor(o4)
            a3 = M.deterministic_actions(l3, 1).to(torch.int32)
            a4 = M.deterministic_actions(l4, 2).to(torch.int32)
            if idx is None:
                act[:, 0, :4] = a3
                act[:, 1, :3] = a4
            else:
                act[idx, 0, :4] = a3
                act[idx, 1, :3] = a4
            if return_logits:
                logits_out[k] = (idx, l3, l4)
        return (act, logits_out) if return_logits else act
