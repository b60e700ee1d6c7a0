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


def generate(name, level, mode, arena, n_steps, kw):
    env = rh.ReferenceEnv(rh.make_namespace(level=level, agent_mode=mode, **kw), SEED, arena)
    rng = np.random.default_rng(arena + 100 * level)
    actions = action_stream(rng, n_steps)
    o1, o2 = env.reset()
    rec = {k: [] for k in ("obs1", "obs2", "rew", "present", "done", "scalars") + F64 + I32}
    reset_obs1, reset_obs2 = [o1], [o2]
    for t in range(n_steps):
        o1, o2, r, pres, done = env.step(actions[t])
        st = env.state()
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
    print(f"{name}: {n_steps} steps, {len(reset_obs1) - 1} episodes, "
          f"kills(agents alive min)={out['scalars'][:, 1].min()} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    for c in CASES:
        generate(*c)
