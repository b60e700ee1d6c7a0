"""Drop-in for the reference's envs/env_hetero.py::LowLevelEnv (one arena, dict API).

Same constructor (`LowLevelEnv(env_config={"args": Namespace})`), attributes and return
shapes as the reference (env_hetero.py:20-60, env_base.py:79-109), so an RLlib-style driver
(`PPOConfig.environment(env=LowLevelEnv, env_config=...)`, train_hetero.py:215) can construct
and step it unchanged.  Internally it is a 1-arena VecLowLevelEnv stepped through the host
entry points of the C ABI -- the CUDA kernel is the only implementation.
"""
from __future__ import annotations

import numpy as np

from .spaces import Box, Dict, MultiDiscrete
from .vec_env import VecLowLevelEnv

try:  # pragma: no cover - ray is absent in the build image
    from ray.rllib.env.multi_agent_env import MultiAgentEnv as _Base  # type: ignore
except Exception:  # noqa: BLE001
    class _Base:  # fallback base so the class imports without ray
        def __init__(self):
            pass

ACTION_DIM_AC1, ACTION_DIM_AC2 = 4, 3
OBS_AC1, OBS_AC2, OBS_ESC_AC1, OBS_ESC_AC2 = 26, 24, 30, 29


class LowLevelEnv(_Base):
    def __init__(self, env_config):
        self.args = env_config.get("args", None)
        self.agent_mode = self.args.agent_mode
        self.opp_mode = "fight"
        self.obs_fight = {1: OBS_AC1, 2: OBS_AC2, 3: OBS_AC1, 4: OBS_AC2}
        self.obs_esc = {1: OBS_ESC_AC1, 2: OBS_ESC_AC2, 3: OBS_ESC_AC1, 4: OBS_ESC_AC2}
        self.obs_dim_map = self.obs_fight if self.agent_mode == "fight" else self.obs_esc
        self._obs_space_in_preferred_format = True
        self.observation_space = Dict({i: Box(np.zeros(self.obs_dim_map[i]), np.ones(self.obs_dim_map[i]),
                                              dtype=np.float32) for i in range(1, 5)})
        self._action_space_in_preferred_format = True
        self.action_space = Dict({1: MultiDiscrete([13, 9, 2, 2]), 2: MultiDiscrete([13, 9, 2]),
                                  3: MultiDiscrete([13, 9, 2, 2]), 4: MultiDiscrete([13, 9, 2])})
        self._agent_ids = set(range(1, self.args.num_agents + 1))
        self._skip_env_checking = True
        self.steps = 0
        self.rewards = {}
        seed = int(env_config.get("seed", 0))
        arena_id = int(env_config.get("arena_id", env_config.get("worker_index", 0)
                                      if hasattr(env_config, "get") else 0))
        self._vec = VecLowLevelEnv(1, self.args, device=int(env_config.get("device", 0)), seed=seed,
                                   arena_base=arena_id, autoreset=False,
                                   opponent_policies=env_config.get("opponent_policies"),
                                   allow_standin_opponents=bool(env_config.get("allow_standin_opponents", False)))
        self._alive_at_step_start = np.array([1, 1])
        super().__init__()

    # env_hetero.py:53-60
    def reset(self, *, seed=None, options=None):
        o1, o2 = self._vec.reset_host()
        self.steps = 0
        self._alive_at_step_start = np.array([1, 1])
        return {1: o1[0].copy(), 2: o2[0].copy()}, {}

    # env_base.py:79-109
    def step(self, action):
        self.rewards = {}
        if action:
            act = np.zeros((1, 2, 4), np.int32)
            for aid in (1, 2):
                if aid in action:
                    a = np.asarray(action[aid]).reshape(-1)
                    nvec = (13, 9, 2, 2) if aid == 1 else (13, 9, 2)
                    if len(a) < len(nvec) or any(int(v) < 0 or int(v) >= n for v, n in zip(a, nvec)):
                        raise ValueError(f"action for agent {aid} outside MultiDiscrete{list(nvec)}: {a}")
                    act[0, aid - 1, :len(nvec)] = a[:len(nvec)]
            o1, o2, r, d = self._vec.step_host(act)
            self.steps += 1
            st = self._vec.get_state()
            if int(st["error"][0]) != 0:
                # the reference raises Exception from set_heading / set_speed (ac1.py:58-67)
                raise ValueError(f"unit heading/speed out of range (error bits {int(st['error'][0])})")
            for k, aid in enumerate((1, 2)):
                if self._alive_at_step_start[k]:
                    self.rewards[aid] = float(r[0, k])
            self._alive_at_step_start = st["alive"][0, :2].copy()
            done = bool(d[0])
            self._last_obs = {1: o1[0].copy(), 2: o2[0].copy()}
        else:
            done = self._done_flag() if hasattr(self, "_last_obs") else False
        terminateds = truncateds = {}
        truncateds["__all__"] = terminateds["__all__"] = done
        return self._last_obs, self.rewards, terminateds, truncateds, {}

    def _done_flag(self):
        st = self._vec.get_state()
        return bool(st["alive_agents"][0] <= 0 or st["alive_opps"][0] <= 0 or st["steps"][0] >= self.args.horizon)

    def state(self):
        return self._last_obs
