"""oracle/oracle.py -- TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.

ctypes front-end of oracle/_build/liboracle.so (oracle/geodesic.c + oracle/hhmarl_oracle.c):
the scalar C restatement of the reference's low-level environment
(envs/env_hetero.py, envs/env_base.py, warsim/simulator/*) and of the third-party
geographiclib==2.0 geodesic it calls.  Only tests/, __graft_entry__.smoke() and bench.py's
CPU-baseline legs import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

ORC_MAX_AC = 6
D = ctypes.c_double
I32 = ctypes.c_int32


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (seconds)."""
    srcs = [os.path.join(_HERE, f) for f in
            ("geodesic.c", "hhmarl_oracle.c", "geodesic.h", "hhmarl_oracle.h", "philox.h")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


class OrcArgs(ctypes.Structure):
    _fields_ = [("level", I32), ("agent_mode", I32), ("horizon", I32), ("num_agents", I32),
                ("num_opps", I32), ("esc_dist_rew", I32), ("friendly_kill", I32),
                ("friendly_punish", I32), ("map_size", D), ("rew_scale", D), ("glob_frac", D),
                ("hier_action_assess", I32), ("hier_opp_fight_ratio", I32)]


class OrcState(ctypes.Structure):
    _fields_ = (
        [(n, D * ORC_MAX_AC) for n in ("lat", "lon", "heading", "speed", "new_heading", "new_speed",
                                        "cannon_remain", "cannon_burst", "cannon_max")]
        + [(n, I32 * ORC_MAX_AC) for n in ("missile_remain", "rocket_max", "missile_wait", "alive",
                                            "has_missile", "opp_to_attack", "ac_type")]
        + [(n, D * ORC_MAX_AC) for n in ("r_lat", "r_lon", "r_heading", "r_new_heading", "r_speed")]
        + [(n, I32 * ORC_MAX_AC) for n in ("r_alive", "r_target", "r_id", "r_age")]
        + [(n, I32) for n in ("steps", "alive_agents", "alive_opps", "escaping", "escaping_time",
                              "next_unit_id", "opp_mode", "policy_set", "error")]
        + [("draws_g", ctypes.c_uint64), ("draws_c", ctypes.c_uint64)]
    )


POLICY_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                             ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_int,
                             ctypes.POINTER(I32))

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        P = ctypes.POINTER
        L.orc_geod_direct.argtypes = [D, D, D, D, P(D), P(D), P(D)]
        L.orc_geod_direct.restype = None
        L.orc_geod_inverse.argtypes = [D, D, D, D, P(D), P(D), P(D)]
        L.orc_geod_inverse.restype = None
        L.orc_geodetic_distance_km.argtypes = [D, D, D, D]
        L.orc_geodetic_distance_km.restype = D
        L.orc_geodetic_bearing_deg.argtypes = [D, D, D, D]
        L.orc_geodetic_bearing_deg.restype = D
        L.orc_geod_last_numit.restype = ctypes.c_int
        L.orc_env_create.argtypes = [P(OrcArgs), ctypes.c_uint64, ctypes.c_uint32]
        L.orc_env_create.restype = ctypes.c_void_p
        L.orc_env_destroy.argtypes = [ctypes.c_void_p]
        L.orc_env_destroy.restype = None
        L.orc_env_set_policy_fn.argtypes = [ctypes.c_void_p, POLICY_FN, ctypes.c_void_p]
        L.orc_env_set_policy_fn.restype = None
        L.orc_env_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_env_reset.restype = None
        L.orc_env_step.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 5
        L.orc_env_step.restype = ctypes.c_int
        L.orc_env_get_state.argtypes = [ctypes.c_void_p, P(OrcState)]
        L.orc_env_get_state.restype = None
        L.orc_obs_len.argtypes = [P(OrcArgs), ctypes.c_int]
        L.orc_obs_len.restype = ctypes.c_int
        L.orc_hier_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_hier_reset.restype = None
        L.orc_hier_step.argtypes = [ctypes.c_void_p] * 5
        L.orc_hier_step.restype = ctypes.c_int
        L.orc_hier_eval_info.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_hier_eval_info.restype = None
        L.orc_env_run_random.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64]
        L.orc_env_run_random.restype = ctypes.c_uint64
        _lib = L
    return _lib


# ------------------------------------------------------------------ geodesic
def geod_direct(lat, lon, azi, s12):
    a, b, c = D(), D(), D()
    lib().orc_geod_direct(lat, lon, azi, s12, a, b, c)
    return a.value, b.value, c.value


def geod_inverse(lat1, lon1, lat2, lon2):
    a, b, c = D(), D(), D()
    lib().orc_geod_inverse(lat1, lon1, lat2, lon2, a, b, c)
    return a.value, b.value, c.value


# ------------------------------------------------------------------ RNG contract (python twin of philox.h)
_M0, _M1, _W0, _W1, _MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


class PhiloxStream:
    """One logical stream (G or C) of one arena under the SURVEY.md A.5 contract.

    Duck-types the subset of `random.Random` the reference calls: random, uniform, randint,
    choices (2 outcomes)."""

    def __init__(self, seed: int, arena_id: int, stream: int):
        self.key = (seed & _MASK, (seed >> 32) & _MASK)
        self.arena = arena_id & _MASK
        self.stream = stream
        self.draw = 0

    def random(self) -> float:
        w = philox4x32_10((self.draw & _MASK, (self.draw >> 32) & _MASK, self.arena, self.stream), self.key)
        self.draw += 1
        return ((w[0] >> 5) * 67108864.0 + (w[1] >> 6)) * (1.0 / 9007199254740992.0)

    def uniform(self, a, b) -> float:
        return a + (b - a) * self.random()

    def randint(self, a, b) -> int:
        return a + int(self.random() * (b - a + 1))

    def choices(self, population, weights=None, k=1):
        assert len(population) == 2 and k == 1
        w0, w1 = weights
        return [population[int(self.random() * (w0 + w1) >= w0)]]


# ------------------------------------------------------------------ env
def make_args(level=1, agent_mode="fight", horizon=None, map_size=0.3, rew_scale=1.0, glob_frac=0.0,
              esc_dist_rew=False, friendly_kill=True, friendly_punish=False) -> OrcArgs:
    if horizon is None:
        horizon = {1: 150, 2: 200, 3: 300, 4: 350, 5: 400}[level]  # config.py:95
    return OrcArgs(level=level, agent_mode=0 if agent_mode == "fight" else 1, horizon=horizon,
                   num_agents=2, num_opps=2, esc_dist_rew=int(esc_dist_rew),
                   friendly_kill=int(friendly_kill), friendly_punish=int(friendly_punish),
                   map_size=map_size, rew_scale=rew_scale, glob_frac=glob_frac, hier_action_assess=1,
                   hier_opp_fight_ratio=75)


def make_hier_args(horizon=500, map_size=0.5, rew_scale=1.0, glob_frac=0.0, friendly_kill=True,
                   hier_action_assess=True, hier_opp_fight_ratio=75, level=1, eval_info=False) -> OrcArgs:
    """Config(1) defaults of the reference (config.py:17-57,98): 3-vs-3, map 0.5, horizon 500, level left at 1.
    `eval_info` changes nothing in the env (env_base.py:91 only builds the info dict): see OracleHierEnv.eval_info."""
    return OrcArgs(level=level, agent_mode=0, horizon=horizon, num_agents=3, num_opps=3, esc_dist_rew=0,
                   friendly_kill=int(friendly_kill), friendly_punish=0, map_size=map_size, rew_scale=rew_scale,
                   glob_frac=glob_frac, hier_action_assess=int(hier_action_assess),
                   hier_opp_fight_ratio=int(hier_opp_fight_ratio))


class OracleEnv:
    """Single-arena C oracle env with the vector-env calling convention used by the tests."""

    def __init__(self, args: OrcArgs, seed: int, arena_id: int, policy_fn=None):
        self.args = args
        self._h = lib().orc_env_create(ctypes.byref(args), seed, arena_id)
        self.len1 = lib().orc_obs_len(ctypes.byref(args), 1)
        self.len2 = lib().orc_obs_len(ctypes.byref(args), 2)
        self.obs1 = np.zeros(32, np.float32)
        self.obs2 = np.zeros(32, np.float32)
        self.rew = np.zeros(2, np.float64)
        self.present = np.zeros(2, np.int32)
        self._cb = None
        if policy_fn is not None:
            self.set_policy_fn(policy_fn)

    def set_policy_fn(self, fn):
        def _tramp(user, unit_id, ac_type, mode, pset, obs_p, obs_len, act_p):
            obs = np.ctypeslib.as_array(obs_p, shape=(obs_len,)).copy()
            act = fn(unit_id, ac_type, mode, pset, obs)
            for i, a in enumerate(act):
                act_p[i] = int(a)
        self._cb = POLICY_FN(_tramp)
        lib().orc_env_set_policy_fn(self._h, self._cb, None)

    def reset(self):
        lib().orc_env_reset(self._h, self.obs1.ctypes.data, self.obs2.ctypes.data)
        return self.obs1[:self.len1].copy(), self.obs2[:self.len2].copy()

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32).reshape(2, 4)
        done = lib().orc_env_step(self._h, a.ctypes.data, self.obs1.ctypes.data, self.obs2.ctypes.data,
                                  self.rew.ctypes.data, self.present.ctypes.data)
        return (self.obs1[:self.len1].copy(), self.obs2[:self.len2].copy(), self.rew.copy(),
                self.present.copy(), bool(done))

    def state(self) -> OrcState:
        s = OrcState()
        lib().orc_env_get_state(self._h, ctypes.byref(s))
        return s

    def run_random(self, n_steps: int, action_seed: int = 1) -> int:
        return lib().orc_env_run_random(self._h, n_steps, action_seed)

    def close(self):
        if self._h:
            lib().orc_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# key order of the reference's info dict (env_base.py:104-106)
EVAL_INFO_KEYS = ("agents_win", "opps_win", "draw", "agent_fight", "agent_escape", "opp_fight", "opp_escape",
                  "agent_steps", "opp_steps", "opp1", "opp2", "opp3")


class OracleHierEnv:
    """Single-arena C oracle of the reference's HighLevelEnv (envs/env_hier.py)."""

    def __init__(self, args: OrcArgs, seed: int, arena_id: int, policy_fn):
        self.args = args
        self._h = lib().orc_env_create(ctypes.byref(args), seed, arena_id)
        self.obs = np.zeros((3, 34), np.float32)
        self.rew = np.zeros(3, np.float64)
        self.info = np.zeros(8, np.int32)
        self.set_policy_fn(policy_fn)

    def set_policy_fn(self, fn):
        def _tramp(user, unit_id, ac_type, mode, pset, obs_p, obs_len, act_p):
            obs = np.ctypeslib.as_array(obs_p, shape=(obs_len,)).copy()
            act = fn(unit_id, ac_type, mode, pset, obs)
            for i, a in enumerate(act):
                act_p[i] = int(a)
        self._cb = POLICY_FN(_tramp)
        lib().orc_env_set_policy_fn(self._h, self._cb, None)

    def reset(self):
        lib().orc_hier_reset(self._h, self.obs.ctypes.data)
        return self.obs.copy()

    def step(self, commander_actions):
        a = np.ascontiguousarray(commander_actions, dtype=np.int32)
        done = lib().orc_hier_step(self._h, a.ctypes.data, self.obs.ctypes.data, self.rew.ctypes.data,
                                   self.info.ctypes.data)
        return self.obs.copy(), self.rew.copy(), bool(done), self.info.copy()

    def eval_info(self):
        """env_base.py:91-107 for the step just made, as {key: int} in the reference's key order."""
        out = np.zeros(len(EVAL_INFO_KEYS), np.int32)
        lib().orc_hier_eval_info(self._h, out.ctypes.data)
        return dict(zip(EVAL_INFO_KEYS, (int(v) for v in out)))

    def state(self) -> OrcState:
        s = OrcState()
        lib().orc_env_get_state(self._h, ctypes.byref(s))
        return s

    def close(self):
        if self._h:
            lib().orc_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
