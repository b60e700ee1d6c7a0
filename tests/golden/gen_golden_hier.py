"""tests/golden/gen_golden_hier.py -- golden trajectories of the reference's UNMODIFIED HighLevelEnv
(envs/env_hier.py, 3-vs-3 commander environment) under the stubs / RNG contract of oracle/ref_harness.py.

The frozen low-level policies of the reference are pickles that do not exist in the repository; they are
replaced at the call site (env_base.py:392-396) by a deterministic function of the observation, and every
query (unit, mode, observation) and its answer are recorded so that a replay can feed the same actions."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402

SEED = 20260926
MAXC = 96  # 16 sub-steps x 6 aircraft
CASES = [("hier_default", 0, 90, {}), ("hier_shared", 1, 70, {"glob_frac": 0.3, "hier_action_assess": False,
                                                               "hier_opp_fight_ratio": 40}),
         ("hier_nofk", 2, 70, {"friendly_kill": False, "rew_scale": 2}),
         # args.eval_info (env_base.py:91-107): the info dict of every step, in the reference's key order
         ("hier_evalinfo", 3, 200, {"eval_info": True, "horizon": 170})]


EVAL_KEYS = ("agents_win", "opps_win", "draw", "agent_fight", "agent_escape", "opp_fight", "opp_escape", "agent_steps",
             "opp_steps", "opp1", "opp2", "opp3")


def pseudo_policy(unit_id, ac_type, mode, obs):
    h = int(np.abs(np.round(np.asarray(obs, np.float64), 4)).sum() * 1e4) + 7 * unit_id + 13 * mode
    heads = (13, 9, 2, 2) if ac_type == 1 else (13, 9, 2)
    return [(h // (1 + 3 * k)) % n for k, n in enumerate(heads)]


def generate(name, arena, n_steps, kw):
    calls = []

    def pol(u, t, m, o):
        a = pseudo_policy(u, t, m, o)
        calls.append((u, t, m, o.copy(), a))
        return a

    env = rh.ReferenceHierEnv(rh.make_hier_namespace(**kw), SEED, arena, pol)
    rng = np.random.default_rng(arena)
    rec = {k: [] for k in ("ca", "obs", "rew", "done", "subs", "ca_out", "scalars", "n_calls", "c_unit", "c_type",
                           "c_mode", "c_act", "c_obs")}
    if kw.get("eval_info"):
        rec["eval"] = []
    resets = [env.reset()]
    for t in range(n_steps):
        ca = rng.integers(0, 3, 3)
        calls.clear()
        o, r, d, ns, ca_out = env.step(ca)
        rec["ca"].append(ca); rec["obs"].append(o); rec["rew"].append(r); rec["done"].append(d); rec["subs"].append(ns)
        rec["ca_out"].append([(-1 if ca_out.get(i) is None else ca_out[i]) for i in range(1, 7)])
        rec["scalars"].append(env.scalars())
        if kw.get("eval_info"):
            assert tuple(env.last_info) == EVAL_KEYS
            rec["eval"].append([int(v) for v in env.last_info.values()])
        cu = np.zeros(MAXC, np.int8); ct = np.zeros(MAXC, np.int8); cm = np.zeros(MAXC, np.int8)
        cact = np.zeros((MAXC, 4), np.int8); cobs = np.zeros((MAXC, 30), np.float32)
        for k, (u, ty, m, ob, a) in enumerate(calls):
            cu[k], ct[k], cm[k] = u, ty, m
            cact[k, :len(a)] = a
            cobs[k, :len(ob)] = ob
        rec["n_calls"].append(len(calls)); rec["c_unit"].append(cu); rec["c_type"].append(ct); rec["c_mode"].append(cm)
        rec["c_act"].append(cact); rec["c_obs"].append(cobs)
        if d:
            resets.append(env.reset())
    out = {k: np.asarray(v) for k, v in rec.items()}
    out.update(resets=np.asarray(resets), meta=np.array([SEED, arena], np.int64), kw=np.array(repr(kw)))
    path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
    np.savez_compressed(path, **out)
    print(name, n_steps, "commander steps,", int(out["subs"].sum()), "sub-steps,", len(resets) - 1, "episodes ->",
          os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    for c in CASES:   # `python gen_golden_hier.py hier_evalinfo` regenerates one case
        if len(sys.argv) == 1 or c[0] in sys.argv[1:]:
            generate(*c)
