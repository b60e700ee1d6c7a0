"""Ray-free policy files with the reference's naming, so the curriculum chain L3 -> L4 -> L5 (self-play against
previously trained policies) works end to end.

The reference exports whole-module pickles of RLlib models as `policies/L{level}_AC{i}_{mode}.pt`
(train_hetero.py:98-107) and torch.load()s them in the env (env_base.py:312-347).  Those pickles need ray to
load; here the same file names hold plain state_dicts of hhmarl_2d_b200.models (identical parameter names).
"""
from __future__ import annotations

import os

import torch

from . import models as M


def policy_path(policy_dir: str, level: int, ac: int, mode: str) -> str:
    return os.path.join(policy_dir, f"L{level}_AC{ac}_{mode}.pt")


def save_policies(policy_dir: str, level: int, mode: str, model1, model2):
    """train_hetero.py:101-105: export ac1_policy / ac2_policy of this level."""
    os.makedirs(policy_dir, exist_ok=True)
    for ac, m in ((1, model1), (2, model2)):
        torch.save({"class": type(m).__name__, "state_dict": m.state_dict()}, policy_path(policy_dir, level, ac, mode))


def load_pair(policy_dir: str, level: int, mode: str, device="cpu"):
    """Both policies of one level.  They share ONE shared_layer module (ac_models_hetero.py:22-27), so the two files must
    carry the same shared-layer tensors; files from different runs are refused instead of silently letting AC2's copy
    overwrite AC1's."""
    m1, m2 = M.build_policy_pair("fight" if mode == "fight" else "escape")
    blobs = [torch.load(policy_path(policy_dir, level, ac, mode), map_location=device) for ac in (1, 2)]
    for k in blobs[0]["state_dict"]:
        if k.startswith("shared_layer.") and not torch.equal(blobs[0]["state_dict"][k], blobs[1]["state_dict"][k]):
            raise ValueError(f"{policy_path(policy_dir, level, 1, mode)} and ..._AC2_...: different {k}; the two policies of a "
                             "level share one SHARED_LAYER and must come from the same training run")
    for m, blob in zip((m1, m2), blobs):
        m.load_state_dict(blob["state_dict"])
        m.to(device).eval()
    return m1, m2


def training_state_path(policy_dir: str, level: int, mode: str) -> str:
    return os.path.join(policy_dir, f"L{level}_{mode}_training_state.pt")


def save_training_state(policy_dir: str, level: int, mode: str, learner, sampler=None):
    """What the reference's algo.save() keeps beyond the exported policies (train_hetero.py:281-288): optimiser moments,
    the adaptive KL coefficients, the epoch counter, the learner's shuffling RNG and the sampler's action-RNG counters."""
    os.makedirs(policy_dir, exist_ok=True)
    blob = {"learner": learner.state_dict()}
    if sampler is not None:
        blob["sampler_ctr"] = sampler.ctr.detach().cpu()
    torch.save(blob, training_state_path(policy_dir, level, mode))


def load_training_state(policy_dir: str, level: int, mode: str, learner, sampler=None):
    blob = torch.load(training_state_path(policy_dir, level, mode), map_location=learner.flat.device)
    learner.load_state_dict(blob["learner"])
    if sampler is not None and "sampler_ctr" in blob:
        sampler.ctr.copy_(blob["sampler_ctr"].to(sampler.ctr.device))
        sampler.refresh_policy()
    return learner.epoch


def load_opponent_policies(policy_dir: str, level: int, agent_mode: str, device="cpu"):
    """HHMARLBaseEnv._get_policies("LowLevel") (env_base.py:318-331) with the reference's file selection."""
    if agent_mode == "fight":
        if level == 4:
            f1, f2 = load_pair(policy_dir, 3, "fight", device)
            return {"fight_1": f1, "fight_2": f2}
        out = {}
        for k in (3, 4):
            f1, f2 = load_pair(policy_dir, k, "fight", device)
            out[k] = {"fight_1": f1, "fight_2": f2}
        e1, e2 = load_pair(policy_dir, 3, "escape", device)
        out[5] = {"escape_1": e1, "escape_2": e2}
        return out
    f1, f2 = load_pair(policy_dir, 5, "fight", device)   # escape-vs-L5_fight
    return {"fight_1": f1, "fight_2": f2}


def load_highlevel_policies(policy_dir: str, eval_level_ag: int = 5, eval_level_opp: int = 5, eval_hl: bool = True,
                            device="cpu"):
    """HHMARLBaseEnv._get_policies("HighLevel") (env_base.py:332-346): the container VecHighLevelEnv takes as
    `lowlevel_policies`.  Fight policies of level `eval_level_ag`; escape policies trained against L5 fight if
    present, else the L3 ones; in the low-level evaluation mode (eval_hl False) the opponents fight with the
    policies of level `eval_level_opp` ("fight_1_opp" / "fight_2_opp")."""
    out = {}
    out["fight_1"], out["fight_2"] = load_pair(policy_dir, eval_level_ag, "fight", device)
    esc_level = 5 if all(os.path.exists(policy_path(policy_dir, 5, ac, "escape")) for ac in (1, 2)) else 3
    out["escape_1"], out["escape_2"] = load_pair(policy_dir, esc_level, "escape", device)
    if not eval_hl:
        out["fight_1_opp"], out["fight_2_opp"] = load_pair(policy_dir, eval_level_opp, "fight", device)
    return out
