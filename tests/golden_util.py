"""Shared replay helper: drive any env with the vector calling convention through a golden file."""
import ast
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F64 = ("lat", "lon", "heading", "speed", "new_heading", "new_speed", "cannon_remain", "cannon_burst",
       "cannon_max")
R64 = ("r_lat", "r_lon", "r_heading", "r_new_heading", "r_speed")
I32 = ("missile_remain", "rocket_max", "missile_wait", "alive", "has_missile", "opp_to_attack")
RI32 = ("r_alive", "r_target", "r_age")


def golden_files(policy_levels=None):
    """Golden trajectories of the unmodified reference (gen_golden.py).  policy_levels: None = all, False = levels 1-3
    (scripted opponents), True = levels 4/5 (frozen-policy opponents, every policy query recorded)."""
    files = sorted(glob.glob(os.path.join(GOLDEN_DIR, "lowlevel_*.npz")))
    if policy_levels is None:
        return files
    is_pol = lambda f: os.path.basename(f).startswith(("lowlevel_L4", "lowlevel_L5"))  # noqa: E731
    return [f for f in files if is_pol(f) == bool(policy_levels)]


class GoldenPolicy:
    """Replays the recorded policy queries of a level-4/5 golden: `fn` has the oracle's callback signature, checks every
    query (unit, aircraft type, mode, policy set, observation) against the recording and answers with the recorded action."""

    def __init__(self, g, obs_atol=1e-6):
        self.g, self.t, self.k, self.atol = g, 0, 0, obs_atol

    def begin_step(self, t):
        self.t, self.k = t, 0

    def fn(self, unit_id, ac_type, pmode, pset, obs):
        g, t, k = self.g, self.t, self.k
        assert k < int(g["n_calls"][t]), (t, k, "more policy queries than the reference made")
        assert (unit_id, ac_type, pmode, pset) == (int(g["c_unit"][t][k]), int(g["c_type"][t][k]), int(g["c_mode"][t][k]),
                                                  int(g["c_pset"][t][k])), (t, k)
        obs = np.asarray(obs, np.float32)
        np.testing.assert_allclose(obs, g["c_obs"][t][k][:len(obs)], atol=self.atol, err_msg=f"policy query obs t={t} unit={unit_id}")
        assert not g["c_obs"][t][k][len(obs):].any()
        self.k += 1
        return g["c_act"][t][k][:4 if ac_type == 1 else 3].astype(np.int32)

    def end_step(self):
        assert self.k == int(self.g["n_calls"][self.t]), (self.t, self.k, "fewer policy queries than the reference made")


def load(path):
    g = dict(np.load(path))
    seed, arena, level, mode = (int(x) for x in g["meta"])
    kw = ast.literal_eval(str(g["kw"]))
    return g, seed, arena, level, ("fight" if mode == 0 else "escape"), kw


def rel_close(a, b, rtol, atol):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= atol + rtol * np.abs(b))
