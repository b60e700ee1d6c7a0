"""The `args` namespace consumed by the env (mirrors the fields of the reference's config.py:14-56
that envs/env_hetero.py and envs/env_base.py read)."""
from argparse import Namespace

HORIZON_BY_LEVEL = {1: 150, 2: 200, 3: 300, 4: 350, 5: 400}  # config.py:95


def make_args(level=1, agent_mode="fight", horizon=None, num_agents=2, num_opps=2, map_size=0.3,
              rew_scale=1, glob_frac=0.0, esc_dist_rew=False, friendly_kill=True, friendly_punish=False,
              **extra) -> Namespace:
    if horizon is None:
        horizon = HORIZON_BY_LEVEL[level]
    ns = Namespace(level=level, agent_mode=agent_mode, horizon=horizon, num_agents=num_agents,
                   num_opps=num_opps, total_num=num_agents + num_opps, map_size=map_size,
                   rew_scale=rew_scale, glob_frac=glob_frac, esc_dist_rew=esc_dist_rew,
                   friendly_kill=friendly_kill, friendly_punish=friendly_punish, eval_info=False,
                   eval_hl=False, eval_level_ag=5, eval_level_opp=4, hier_opp_fight_ratio=75,
                   hier_action_assess=True)
    for k, v in extra.items():
        setattr(ns, k, v)
    ns.env_config = {"args": ns}
    return ns
