"""oracle/ref_harness.py -- TEST INFRASTRUCTURE. Runs only where /root/reference exists.

Imports the reference's OWN, UNMODIFIED environment files (envs/env_hetero.py, envs/env_base.py,
warsim/simulator/*, warsim/utils/*) from /root/reference under four stub modules for the
third-party packages that are absent here:

  * ray.rllib.env.multi_agent_env.MultiAgentEnv      -> empty base class
  * gymnasium.spaces.{Dict,Box,MultiDiscrete,Discrete} -> inert containers
  * warsim.scenplotter.scenario_plotter              -> inert names (rendering, out of scope)
  * geographiclib.geodesic.Geodesic.WGS84            -> oracle/geodesic.c (Karney restatement)

and replaces the unseeded RNGs by the shared Philox contract (oracle/philox.h):
the module-global name `random` in envs.env_base, envs.env_hetero and warsim.simulator.ac1, and
`env.sim.rnd_gen` after every reset.  Used to (a) generate tests/golden/*.npz and (b) pin the C
oracle (tests/test_oracle_vs_reference.py).  Nothing here travels to the GPU box.
"""
from __future__ import annotations

import os
import sys
import types
from argparse import Namespace

import numpy as np

REFERENCE_ROOT = os.environ.get("HH_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import oracle as orc  # noqa: E402  (oracle/oracle.py)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "envs", "env_hetero.py"))


# ------------------------------------------------------------------ stubs
class _Geodesic:
    DISTANCE, AZIMUTH, LATITUDE, LONGITUDE = 1, 2, 4, 8

    class _WGS84:
        @staticmethod
        def Inverse(lat1, lon1, lat2, lon2, outmask=None):
            s12, azi1, azi2 = orc.geod_inverse(float(lat1), float(lon1), float(lat2), float(lon2))
            return {"s12": s12, "azi1": azi1, "azi2": azi2}

        @staticmethod
        def Direct(lat1, lon1, azi1, s12, outmask=None):
            lat2, lon2, azi2 = orc.geod_direct(float(lat1), float(lon1), float(azi1), float(s12))
            return {"lat2": lat2, "lon2": lon2, "azi2": azi2}

    WGS84 = _WGS84()


class _Inert:
    def __init__(self, *a, **k):
        self.args, self.kwargs = a, k

    def __call__(self, *a, **k):
        return None

    def __getattr__(self, name):
        return _Inert()


class _RandomProxy:
    """Stands in for the module `random`; forwards to the current arena's G stream."""

    def __init__(self):
        self.stream = None
        self.Random = None  # cmano_simulator.py:88 calls random.Random(seed); patched separately

    def randint(self, a, b):
        return self.stream.randint(a, b)

    def uniform(self, a, b):
        return self.stream.uniform(a, b)

    def random(self):
        return self.stream.random()

    def choices(self, population, weights=None, k=1):
        return self.stream.choices(population, weights=weights, k=k)


_PROXY = _RandomProxy()
_installed = False


def install():
    """Create the stub modules and import the reference env. Idempotent."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    for n in ("ray", "ray.rllib", "ray.rllib.env"):
        mod(n)
    mae = mod("ray.rllib.env.multi_agent_env")

    class MultiAgentEnv:
        def __init__(self):
            pass

    mae.MultiAgentEnv = MultiAgentEnv

    gym = mod("gymnasium")
    spaces = mod("gymnasium.spaces")
    for n in ("Dict", "Box", "MultiDiscrete", "Discrete"):
        setattr(spaces, n, type(n, (_Inert,), {}))
    gym.spaces = spaces

    mod("geographiclib")
    gg = mod("geographiclib.geodesic")
    gg.Geodesic = _Geodesic

    sys.path.insert(0, REFERENCE_ROOT)
    import warsim  # noqa: F401  (namespace package of the reference)
    sp = mod("warsim.scenplotter")
    sp.__path__ = []
    plot = mod("warsim.scenplotter.scenario_plotter")
    for n in ("PlotConfig", "ColorRGBA", "StatusMessage", "TopLeftMessage", "Airplane", "PolyLine",
              "Drawable", "Waypoint", "Missile", "ScenarioPlotter"):
        setattr(plot, n, type(n, (_Inert,), {}))

    import envs.env_base as env_base
    import envs.env_hetero as env_hetero
    import warsim.simulator.ac1 as ac1

    env_base.random = _PROXY
    env_hetero.random = _PROXY
    ac1.random = _PROXY
    _installed = True


def make_namespace(level=1, agent_mode="fight", horizon=None, map_size=0.3, rew_scale=1, glob_frac=0.0,
                   esc_dist_rew=False, friendly_kill=True, friendly_punish=False) -> Namespace:
    """The `args` namespace the env reads (config.py:14-56,94-107), built without argparse."""
    if horizon is None:
        horizon = {1: 150, 2: 200, 3: 300, 4: 350, 5: 400}[level]
    return Namespace(level=level, horizon=horizon, agent_mode=agent_mode, num_agents=2, num_opps=2,
                     total_num=4, map_size=map_size, rew_scale=rew_scale, glob_frac=glob_frac,
                     esc_dist_rew=esc_dist_rew, friendly_kill=friendly_kill,
                     friendly_punish=friendly_punish, eval_info=False, eval_hl=True,   # config.py:50 default
                     eval_level_ag=5, eval_level_opp=4, hier_opp_fight_ratio=75,
                     hier_action_assess=True)


class ReferenceEnv:
    """The reference's LowLevelEnv driven under the RNG contract for one (seed, arena_id)."""

    def __init__(self, args: Namespace, seed: int, arena_id: int, policy_fn=None):
        """`policy_fn(unit_id, ac_type, mode, policy_set, obs) -> action` (levels 4/5; mode 0 fight / 1 escape; policy_set =
        k of env_hetero.py:57 at level-5 fight, else 0) stands in for the pickled RLlib policies, which are not in the
        repository (env_base.py:312-347): it is called at the point where the reference calls
        self.policy[...](input_dict=..., state=..., seq_lens=...) (env_base.py:392-396) and its answer comes back as
        logits whose per-head argmax (env_base.py:373-382) is that action."""
        install()
        import envs.env_base as env_base
        from envs.env_hetero import LowLevelEnv

        self.args = args
        self.g = orc.PhiloxStream(seed, arena_id, 0)
        self.c = orc.PhiloxStream(seed, arena_id, 1)
        if args.level >= 4:
            if policy_fn is None:
                raise ValueError("levels 4/5 need a policy_fn")
            import torch
            outer = self

            class _Pol:
                def __init__(self, mode, ac_type, pset):
                    self.mode, self.ac_type, self.pset = mode, ac_type, pset

                def __call__(self, input_dict=None, state=None, seq_lens=None):
                    o = input_dict["obs"]
                    assert set(o) == {"obs_1_own", "obs_2", "act_1_own", "act_2"}
                    assert not o["obs_2"].any() and not o["act_1_own"].any() and not o["act_2"].any()
                    obs = o["obs_1_own"][0].numpy()
                    act = policy_fn(outer._cur_unit, self.ac_type, self.mode, self.pset, obs)
                    heads = (13, 9, 2, 2) if self.ac_type == 1 else (13, 9, 2)
                    logits = torch.full((1, sum(heads)), -10.0)
                    k = 0
                    for h, a in zip(heads, act):
                        logits[0, k + int(a)] = 10.0
                        k += h
                    return logits, []

            def _get(this, mode):      # the key structure of env_base.py:318-331
                assert mode == "LowLevel"
                this.policy = {}
                if args.agent_mode == "fight" and args.level == 5:
                    this.policies = {k: ({"fight_1": _Pol(0, 1, k), "fight_2": _Pol(0, 2, k)} if k <= 4 else
                                         {"escape_1": _Pol(1, 1, k), "escape_2": _Pol(1, 2, k)}) for k in (3, 4, 5)}
                elif args.agent_mode == "fight" or args.level == 5:
                    this.policy = {"fight_1": _Pol(0, 1, 0), "fight_2": _Pol(0, 2, 0)}

            orig = env_base.HHMARLBaseEnv._get_policies
            orig_pa = env_base.HHMARLBaseEnv._policy_actions

            def _pa(this, policy_type, agent_id, unit):
                outer._cur_unit = agent_id
                return orig_pa(this, policy_type, agent_id, unit)

            env_base.HHMARLBaseEnv._get_policies = _get
            try:
                self.env = LowLevelEnv({"args": args})
            finally:
                env_base.HHMARLBaseEnv._get_policies = orig
            self.env._policy_actions = types.MethodType(_pa, self.env)
        else:
            self.env = LowLevelEnv({"args": args})

    def _bind(self):
        _PROXY.stream = self.g

    def reset(self):
        self._bind()
        self._units = {}
        # CmanoSimulator(...) builds random.Random(None) from the *real* random module
        # (cmano_simulator.py:9,88); we overwrite it right after, as A.5 specifies.
        import envs.env_base as env_base
        orig_sim = env_base.CmanoSimulator
        c = self.c

        def _sim(*a, **k):
            s = orig_sim(*a, **k)
            s.rnd_gen = c
            return s

        env_base.CmanoSimulator = _sim
        try:
            obs, info = self.env.reset()
        finally:
            env_base.CmanoSimulator = orig_sim
        return obs[1], obs[2]

    def step(self, actions):
        self._bind()
        a = np.asarray(actions).reshape(2, 4)
        act = {1: a[0, :4].copy(), 2: a[1, :3].copy()}
        obs, rew, term, trunc, info = self.env.step(act)
        r = np.array([rew.get(1, 0.0), rew.get(2, 0.0)], np.float64)
        present = np.array([1 in rew, 2 in rew], np.int32)
        return obs[1], obs[2], r, present, bool(term["__all__"])

    def state(self) -> dict:
        """Same fields as oracle.OrcState, read off the reference's objects."""
        e, sim = self.env, self.env.sim
        out = {k: np.zeros(4) for k in ("lat", "lon", "heading", "speed", "new_heading", "new_speed",
                                         "cannon_remain", "cannon_burst", "cannon_max",
                                         "r_lat", "r_lon", "r_heading", "r_new_heading", "r_speed")}
        out.update({k: np.zeros(4, np.int32) for k in ("missile_remain", "rocket_max", "missile_wait",
                                                        "alive", "has_missile", "opp_to_attack",
                                                        "r_alive", "r_target", "r_id", "r_age")})
        units = getattr(self, "_units", {})
        units.update(sim.active_units)
        self._units = units
        for i in range(1, 5):
            u = units.get(i)
            if u is None or (u.id != i):
                continue
            k = i - 1
            out["lat"][k], out["lon"][k] = u.position.lat, u.position.lon
            out["heading"][k], out["speed"][k] = u.heading, u.speed
            out["new_heading"][k], out["new_speed"][k] = u.new_heading, u.new_speed
            out["cannon_remain"][k], out["cannon_burst"][k] = u.cannon_remain_secs, u.cannon_current_burst_secs
            out["cannon_max"][k] = u.cannon_max
            out["missile_remain"][k], out["rocket_max"][k] = u.missile_remain, u.rocket_max
            out["missile_wait"][k] = e.missile_wait[i]
            out["alive"][k] = int(sim.unit_exists(i))
            out["has_missile"][k] = int(bool(u.actual_missile))
            out["opp_to_attack"][k] = e.opp_to_attack[i] or 0
            m = u.actual_missile
            if m:
                out["r_lat"][k], out["r_lon"][k] = m.position.lat, m.position.lon
                out["r_heading"][k], out["r_new_heading"][k] = m.heading, m.new_heading
                out["r_speed"][k] = float(m.speed)
                out["r_alive"][k] = int(sim.unit_exists(m.id))
                out["r_target"][k], out["r_id"][k] = m.target.id, m.id
                out["r_age"][k] = (sim.utc_time - m.firing_time).seconds
        out["scalars"] = np.array([e.steps, e.alive_agents, e.alive_opps, int(e.hardcoded_opps_escaping),
                                   e.opps_escaping_time, sim._next_unit_id, self.g.draw, self.c.draw],
                                  np.int64)
        return out



# ------------------------------------------------------------------ reference MODELS under stubs
_models_installed = False


def install_model_stubs():
    """Stub the ray.rllib symbols imported by models/ac_models_hetero.py:1-9 so that the reference's
    model file can be imported UNMODIFIED.  SlimFC / add_time_dimension restate RLlib 2.4 (absent here,
    SURVEY.md Appendix C -> that part of the parity stays unpinned); the network structure, slicing and
    concatenation order come from the reference's own code."""
    global _models_installed
    if _models_installed:
        return
    install()
    import torch
    import torch.nn as nn
    sys.path.insert(0, os.path.dirname(_HERE))
    from hhmarl_2d_b200.models import SlimFC, add_time_dimension

    def mod(name):
        m = sys.modules.get(name) or types.ModuleType(name)
        sys.modules[name] = m
        return m

    for n in ("ray.rllib.models", "ray.rllib.models.torch", "ray.rllib.utils", "ray.rllib.policy"):
        mod(n)

    class ModelV2:
        pass

    class TorchModelV2(ModelV2):
        def __init__(self, obs_space, action_space, num_outputs, model_config, name):
            self.obs_space, self.action_space, self.num_outputs = obs_space, action_space, num_outputs
            self.model_config, self.name = model_config, name

    class RecurrentNetwork(TorchModelV2):
        pass

    mod("ray.rllib.models.modelv2").ModelV2 = ModelV2
    mod("ray.rllib.models.torch.misc").SlimFC = SlimFC
    mod("ray.rllib.models.torch.torch_modelv2").TorchModelV2 = TorchModelV2
    mod("ray.rllib.models.torch.recurrent_net").RecurrentNetwork = RecurrentNetwork
    mod("ray.rllib.utils.annotations").override = lambda cls: (lambda f: f)
    mod("ray.rllib.utils.framework").try_import_torch = lambda: (torch, nn)

    def _atd(padded_inputs, seq_lens=None, framework="torch", time_major=False, **kw):
        return add_time_dimension(padded_inputs, seq_lens)

    mod("ray.rllib.policy.rnn_sequencing").add_time_dimension = _atd
    _models_installed = True


def reference_models():
    """{'Fight1': cls, ...} of the reference's own classes + its module (for SHARED_LAYER)."""
    install_model_stubs()
    import models.ac_models_hetero as m
    return m


# ------------------------------------------------------------------ reference HighLevelEnv under stubs
def make_hier_namespace(horizon=500, map_size=0.5, rew_scale=1, glob_frac=0.0, friendly_kill=True,
                        hier_action_assess=True, hier_opp_fight_ratio=75, level=1, eval_info=False) -> Namespace:
    """Config(1) of the reference (config.py:17-57, 94-107) without argparse."""
    return Namespace(level=level, horizon=horizon, agent_mode="fight", num_agents=3, num_opps=3, total_num=6,
                     map_size=map_size, rew_scale=rew_scale, glob_frac=glob_frac, esc_dist_rew=False,
                     friendly_kill=friendly_kill, friendly_punish=False, eval_info=eval_info, eval_hl=True,
                     eval_level_ag=5, eval_level_opp=5, hier_opp_fight_ratio=hier_opp_fight_ratio,
                     hier_action_assess=hier_action_assess)


class ReferenceHierEnv:
    """The reference's HighLevelEnv under the RNG contract.  `policy_fn(unit_id, ac_type, mode, obs) -> action`
    replaces the pickled RLlib policies (env_base.py:312-347) at the point where the reference calls
    self.policy[...](input_dict=..., state=..., seq_lens=...) (env_base.py:392-396)."""

    def __init__(self, args: Namespace, seed: int, arena_id: int, policy_fn):
        install()
        import torch
        import envs.env_base as env_base
        import envs.env_hier as env_hier
        env_hier.random = _PROXY
        self.g = orc.PhiloxStream(seed, arena_id, 0)
        self.c = orc.PhiloxStream(seed, arena_id, 1)
        outer = self

        class _Pol:
            def __init__(self, mode, ac_type):
                self.mode, self.ac_type = mode, ac_type

            def __call__(self, input_dict=None, state=None, seq_lens=None):
                obs = input_dict["obs"]["obs_1_own"][0].numpy()
                act = policy_fn(outer._cur_unit, self.ac_type, self.mode, obs)
                heads = (13, 9, 2, 2) if self.ac_type == 1 else (13, 9, 2)
                logits = torch.full((1, sum(heads)), -10.0)
                o = 0
                for h, a in zip(heads, act):
                    logits[0, o + int(a)] = 10.0
                    o += h
                return logits, []

        def _get(this, mode):
            this.policy = {f"{m}_{t}": _Pol(0 if m == "fight" else 1, t) for m in ("fight", "escape") for t in (1, 2)}

        orig_get = env_base.HHMARLBaseEnv._get_policies
        orig_pa = env_base.HHMARLBaseEnv._policy_actions

        def _pa(this, policy_type, agent_id, unit):
            outer._cur_unit = agent_id
            return orig_pa(this, policy_type, agent_id, unit)

        env_base.HHMARLBaseEnv._get_policies = _get
        try:
            self.env = env_hier.HighLevelEnv({"args": args})
        finally:
            env_base.HHMARLBaseEnv._get_policies = orig_get
        self.env._policy_actions = types.MethodType(_pa, self.env)

    def reset(self):
        _PROXY.stream = self.g
        import envs.env_base as env_base
        orig_sim = env_base.CmanoSimulator
        c = self.c

        def _sim(*a, **k):
            s = orig_sim(*a, **k)
            s.rnd_gen = c
            return s

        env_base.CmanoSimulator = _sim
        try:
            obs, _ = self.env.reset()
        finally:
            env_base.CmanoSimulator = orig_sim
        return np.stack([obs[i] for i in (1, 2, 3)])

    def step(self, commander_actions):
        _PROXY.stream = self.g
        ca = {i + 1: int(a) for i, a in enumerate(commander_actions)}
        steps0 = self.env.steps
        obs, rew, term, trunc, info = self.env.step(ca)
        self.last_info = info   # {} unless args.eval_info (env_base.py:91-107)
        return (np.stack([obs[i] for i in (1, 2, 3)]), np.array([rew[i] for i in (1, 2, 3)], np.float64),
                bool(term["__all__"]), self.env.steps - steps0, dict(ca))

    def scalars(self):
        e = self.env
        return [e.steps, e.alive_agents, e.alive_opps, e.sim._next_unit_id, self.g.draw, self.c.draw]
