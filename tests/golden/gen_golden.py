"""tests/golden/gen_golden.py -- generates the committed golden trajectories.

Runs the reference's UNMODIFIED LowLevelEnv (from /root/reference, under the stubs and the
Philox RNG contract of oracle/ref_harness.py) on seeded random MultiDiscrete action streams and
dumps, per step, everything the parity tests compare: observations, rewards, reward-dict
membership, done, the full unit state and the RNG draw counters.  /root/reference does not
exist on the GPU box, so these .npz files are what travels.

    python tests/golden/gen_golden.py        # rewrites tests/golden/lowlevel_*.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402

SEED = 20260925
CASES = [
    # name, level, mode, arena_id, n_steps, extra kwargs
    ("L1_fight", 1, "fight", 0, 450, {}),
    ("L2_fight", 2, "fight", 1, 500, {}),
    ("L3_fight", 3, "fight", 2, 700, {}),
    ("L3_fight_b", 3, "fight", 3, 700, {}),
    ("L3_escape", 3, "escape", 4, 500, {"esc_dist_rew": True}),
    ("L3_fight_shared", 3, "fight", 5, 400, {"glob_frac": 0.5, "friendly_punish": True, "rew_scale": 2}),
    ("L2_fight_nofk", 2, "fight", 6, 400, {"friendly_kill": False}),
]
# levels 4/5: the frozen opponent policies are replaced at the reference's call site by a deterministic function of the
# query (ref_harness.ReferenceEnv policy_fn); every query and its answer are recorded, <= 2 per step (units 3, 4)
POLICY_CASES = [
    ("L4_fight", 4, "fight", 7, 800, {}),
    ("L5_fight", 5, "fight", 8, 1400, {}),                      # exercises policy sets k = 3, 4, 5 (env_hetero.py:55-59)
    ("L5_escape", 5, "escape", 9, 900, {"esc_dist_rew": True}),
]
F64 = ("lat", "lon", "heading", "speed", "new_heading", "new_speed", "cannon_remain", "cannon_burst",
       "cannon_max", "r_lat", "r_lon", "r_heading", "r_new_heading", "r_speed")
I32 = ("missile_remain", "rocket_max", "missile_wait", "alive", "has_missile", "opp_to_attack",
       "r_alive", "r_target", "r_id", "r_age")


def action_stream(rng, n):
    a = np.zeros((n, 2, 4), np.int32)
    a[..., 0] = rng.integers(0, 13, (n, 2))
    a[..., 1] = rng.integers(0, 9, (n, 2))
    a[..., 2] = rng.integers(0, 2, (n, 2))
    a[..., 3] = rng.integers(0, 2, (n, 2))
    # bias towards firing so that cannon/missile bookkeeping is exercised
    return a


def pseudo_policy(unit_id, ac_type, mode, pset, obs):
    """Deterministic stand-in for a frozen policy: the action is a hash of the query (so a wrong observation, unit,
    mode or policy set anywhere changes the trajectory)."""
    h = int(np.abs(np.round(np.asarray(obs, np.float64), 4)).sum() * 1e4) + 7 * unit_id + 13 * mode + 29 * pset
    heads = (13, 9, 2, 2) if ac_type == 1 else (13, 9, 2)
    return [(h // (1 + 3 * k)) % n for k, n in enumerate(heads)]


def generate(name, level, mode, arena, n_steps, kw):
    calls = []

    def pol(u, t, m, ps, o):
        a = pseudo_policy(u, t, m, ps, o)
        calls.append((u, t, m, ps, np.array(o, np.float32), a))
        return a

    env = rh.ReferenceEnv(rh.make_namespace(level=level, agent_mode=mode, **kw), SEED, arena,
                          policy_fn=pol if level >= 4 else None)
    rng = np.random.default_rng(arena + 100 * level)
    actions = action_stream(rng, n_steps)
    o1, o2 = env.reset()
    rec = {k: [] for k in ("obs1", "obs2", "rew", "present", "done", "scalars") + F64 + I32}
    if level >= 4:
        rec.update({k: [] for k in ("n_calls", "c_unit", "c_type", "c_mode", "c_pset", "c_act", "c_obs")})
    reset_obs1, reset_obs2 = [o1], [o2]
    for t in range(n_steps):
        calls.clear()
        o1, o2, r, pres, done = env.step(actions[t])
        st = env.state()
        if level >= 4:
            cu = np.zeros(2, np.int8); ct = np.zeros(2, np.int8); cm = np.zeros(2, np.int8); cp = np.zeros(2, np.int8)
            cact = np.zeros((2, 4), np.int8); cobs = np.zeros((2, 30), np.float32)
            for k, (u, ty, m, ps, ob, a) in enumerate(calls):
                cu[k], ct[k], cm[k], cp[k] = u, ty, m, ps
                cact[k, :len(a)] = a
                cobs[k, :len(ob)] = ob
            rec["n_calls"].append(len(calls)); rec["c_unit"].append(cu); rec["c_type"].append(ct); rec["c_mode"].append(cm)
            rec["c_pset"].append(cp); rec["c_act"].append(cact); rec["c_obs"].append(cobs)
        rec["obs1"].append(o1); rec["obs2"].append(o2); rec["rew"].append(r)
        rec["present"].append(pres); rec["done"].append(done); rec["scalars"].append(st["scalars"])
        for k in F64 + I32:
            rec[k].append(st[k])
        if done:
            o1, o2 = env.reset()
            reset_obs1.append(o1); reset_obs2.append(o2)
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["done"] = out["done"].astype(np.uint8)
    out["present"] = out["present"].astype(np.uint8)
    for k in I32:
        out[k] = out[k].astype(np.int16)
    out.update(actions=actions.astype(np.int8), reset_obs1=np.asarray(reset_obs1),
               reset_obs2=np.asarray(reset_obs2),
               meta=np.array([SEED, arena, level, 0 if mode == "fight" else 1], np.int64),
               kw=np.array(repr(kw)))
    path = os.path.join(ROOT, "tests", "golden", f"lowlevel_{name}.npz")
    np.savez_compressed(path, **out)
    extra = ""
    if level >= 4:
        extra = f", {int(out['n_calls'].sum())} policy queries, policy sets {sorted(set(out['c_pset'][out['c_unit'] > 0].tolist()))}"
    print(f"{name}: {n_steps} steps, {len(reset_obs1) - 1} episodes, "
          f"kills(agents alive min)={out['scalars'][:, 1].min()}{extra} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    for c in CASES + POLICY_CASES:   # `python gen_golden.py L5_fight` regenerates one case
        if len(sys.argv) == 1 or c[0] in sys.argv[1:]:
            generate(*c)
