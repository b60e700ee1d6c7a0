"""VecLowLevelEnv -- N independent 2-vs-2 low-level arenas advanced in lock-step by one fused
sm_100a kernel per step (hhmarl_2d_b200/csrc/hh_api.cu) through the C ABI
(include/hhmarl_b200.h).  Mirrors envs/env_hetero.py::LowLevelEnv (reset / step /
observation_space / action_space) with a leading arena axis.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native as nat
from .config import make_args
from .spaces import Box, Dict, MultiDiscrete

ACTION_NVEC_AC1 = (13, 9, 2, 2)  # env_hetero.py:38
ACTION_NVEC_AC2 = (13, 9, 2)     # env_hetero.py:39


class VecLowLevelEnv:
    """Batched twin of the reference's LowLevelEnv.

    Parameters
    ----------
    n_arenas : number of independent 2-vs-2 arenas.
    args     : namespace with the reference's `args` fields (see hhmarl_2d_b200.make_args).
    device   : CUDA device index.
    seed     : Philox key; arena k uses counter-based streams keyed by (seed, arena_base + k), so
               results do not depend on how arenas are sharded over GPUs.
    autoreset: reset finished arenas inside the step launch (sampler semantics); the returned obs
               of such an arena is the first observation of its next episode.
    """

    def __init__(self, n_arenas: int, args=None, device: int = 0, seed: int = 0, arena_base: int = 0,
                 autoreset: bool = True, opponent_policies=None, allow_standin_opponents: bool = False):
        """opponent_policies (levels 4/5): the frozen self-play policies of `_get_policies("LowLevel")`
        (env_base.py:312-331; checkpoint.load_opponent_policies).  Like the reference, which fails when the policy files
        are missing, a level-4/5 env without them is an error -- unless `allow_standin_opponents` asks for seeded
        random-weight stand-ins (tests and benchmarks: there are no trained weights in the repository)."""
        self.args = args if args is not None else make_args()
        a = self.args
        if a.num_agents != 2 or a.num_opps != 2:
            raise ValueError("VecLowLevelEnv implements the 2-vs-2 low-level scenario")
        if a.level >= 4 and opponent_policies is None and not allow_standin_opponents:
            raise ValueError(f"level {a.level} needs the frozen opponent policies (opponent_policies=..., see "
                             "checkpoint.load_opponent_policies); pass allow_standin_opponents=True for seeded "
                             "random-weight stand-ins")
        self.n_arenas = int(n_arenas)
        self.device_index = int(device)
        cfg = nat.HHConfig(level=a.level, agent_mode=0 if a.agent_mode == "fight" else 1, horizon=a.horizon,
                           esc_dist_rew=int(bool(a.esc_dist_rew)), friendly_kill=int(bool(a.friendly_kill)),
                           friendly_punish=int(bool(a.friendly_punish)), autoreset=int(bool(autoreset)),
                           reserved=0, map_size=float(a.map_size), rew_scale=float(a.rew_scale),
                           glob_frac=float(a.glob_frac), seed=int(seed), arena_base=int(arena_base))
        self._cfg = cfg
        self._h = nat.VP()
        nat.check(nat.lib().hh_create(ctypes.byref(cfg), self.n_arenas, self.device_index, ctypes.byref(self._h)),
                  "hh_create")
        self.obs_dim = (nat.lib().hh_obs_dim(self._h, 1), nat.lib().hh_obs_dim(self._h, 2))
        # reference spaces (env_hetero.py:29-44) for agents 1, 2
        self.observation_space = Dict({1: Box(0.0, 1.0, (self.obs_dim[0],), np.float32),
                                       2: Box(0.0, 1.0, (self.obs_dim[1],), np.float32)})
        self.action_space = Dict({1: MultiDiscrete(ACTION_NVEC_AC1), 2: MultiDiscrete(ACTION_NVEC_AC2)})
        self._agent_ids = {1, 2}
        self._torch = None
        self._bufs = None
        self._opp = None
        self._opp_policies_arg = opponent_policies
        # optional override of the opponents' controller at levels 4/5: fn(dict(obs3 [N,30], obs4 [N,29], pset [N] u8)) ->
        # int32 [N,2,4] CUDA tensor (the golden-replay tests inject the reference's recorded actions through it)
        self.opponent_action_fn = None
        self.fused_opponents = True     # levels 4/5: frozen actors through csrc/hh_policy.cu (False: per-layer torch forward)
        self.opponent_precision = 2     # 2: tcgen05 path (fp32-equivalent logits before the argmax), 0: 3xTF32 on mma.sync, 1: plain TF32
        self.level = int(a.level)

    # ------------------------------------------------------------------ device (torch) API
    def _ensure_torch(self):
        if self._bufs is None:
            import torch
            self._torch = torch
            dev = torch.device("cuda", self.device_index)
            n = self.n_arenas
            self._bufs = dict(obs1=torch.empty((n, self.obs_dim[0]), dtype=torch.float32, device=dev),
                              obs2=torch.empty((n, self.obs_dim[1]), dtype=torch.float32, device=dev),
                              rew=torch.empty((n, 2), dtype=torch.float32, device=dev),
                              done=torch.empty((n,), dtype=torch.uint8, device=dev))
        return self._bufs

    def _stream(self):
        return self._torch.cuda.current_stream(self.device_index).cuda_stream

    def reset(self, mask=None, out=None):
        """Reset all arenas (or those with mask != 0). Returns (obs1 [N,d1], obs2 [N,d2]) CUDA tensors."""
        b = out if out is not None else self._ensure_torch()
        self._ensure_torch()
        mptr = None
        if mask is not None:
            assert mask.is_cuda and mask.dtype == self._torch.uint8 and mask.numel() == self.n_arenas
            mptr = mask.data_ptr()
        nat.check(nat.lib().hh_reset(self._h, mptr, b["obs1"].data_ptr(), b["obs2"].data_ptr(), self._stream()),
                  "hh_reset")
        return b["obs1"], b["obs2"]

    def step(self, actions, out=None):
        """actions: int32 CUDA tensor [N, 2, 4]. Returns (obs1, obs2, rew [N,2], done [N] u8).

        The returned tensors are the env's own output buffers (overwritten by the next call)
        unless `out` (a dict with the same keys) is given.  Enqueued on the current torch stream;
        no host synchronisation."""
        b = out if out is not None else self._ensure_torch()
        self._ensure_torch()
        t = self._torch
        if not (actions.is_cuda and actions.dtype == t.int32 and actions.is_contiguous()
                and actions.numel() == self.n_arenas * 8):
            raise ValueError("actions must be a contiguous int32 CUDA tensor of shape [N, 2, 4]")
        if self.level >= 4:
            opp_actions = self.opponent_actions(actions)
            nat.check(nat.lib().hh_step_finish(self._h, opp_actions.data_ptr(), b["obs1"].data_ptr(),
                                               b["obs2"].data_ptr(), b["rew"].data_ptr(), b["done"].data_ptr(),
                                               self._stream()), "hh_step_finish")
            return b["obs1"], b["obs2"], b["rew"], b["done"]
        nat.check(nat.lib().hh_step(self._h, actions.data_ptr(), b["obs1"].data_ptr(), b["obs2"].data_ptr(),
                                    b["rew"].data_ptr(), b["done"].data_ptr(), self._stream()), "hh_step")
        return b["obs1"], b["obs2"], b["rew"], b["done"]

    def step_range(self, first, count, actions, out=None, central=None):
        """step() for arenas [first, first + count) only (levels 1-3; `first` a multiple of 32).  `actions` and the output tensors
        are the WHOLE batch's ([N, ...]); only the range's rows are read / written.  Two ranges may be in flight on two streams.
        `central` = (flat1_next, flat2_next): the kernel also writes the new observations as the central-critic rows of both
        policies ([N, 7 + d1 + d2] each, columns 7.. ; hh_step_range_central) -- no separate packing launch."""
        b = out if out is not None else self._ensure_torch()
        self._ensure_torch()
        t = self._torch
        if not (actions.is_cuda and actions.dtype == t.int32 and actions.is_contiguous()
                and actions.numel() == self.n_arenas * 8):
            raise ValueError("actions must be a contiguous int32 CUDA tensor of shape [N, 2, 4]")
        if central is not None:
            c1, c2 = central
            d1, d2 = self.obs_dim
            if not (c1.is_cuda and c2.is_cuda and c1.dtype == t.float32 and c2.dtype == t.float32 and c1.shape == c2.shape
                    and c1.shape[0] == self.n_arenas and c1.shape[1] >= 7 + d1 + d2 and c1.stride(1) == 1 and c2.stride(1) == 1
                    and c1.stride(0) == c2.stride(0)):
                raise ValueError("central must be two float32 CUDA tensors [N, >= 7 + d1 + d2] with the same row stride")
            nat.check(nat.lib().hh_step_range_central(self._h, int(first), int(count), actions.data_ptr(), b["obs1"].data_ptr(),
                                                      b["obs2"].data_ptr(), b["rew"].data_ptr(), b["done"].data_ptr(), c1.data_ptr(),
                                                      c2.data_ptr(), int(c1.stride(0)), self._stream()), "hh_step_range_central")
            return
        nat.check(nat.lib().hh_step_range(self._h, int(first), int(count), actions.data_ptr(), b["obs1"].data_ptr(),
                                          b["obs2"].data_ptr(), b["rew"].data_ptr(), b["done"].data_ptr(), self._stream()),
                  "hh_step_range")

    # ------------------------------------------------------------------ levels 4/5: frozen-policy opponents
    def _ensure_opp_bufs(self):
        if getattr(self, "_opp_bufs", None) is None:
            t = self._torch
            dev = t.device("cuda", self.device_index)
            n = self.n_arenas
            self._opp_bufs = dict(obs3=t.empty((n, 30), dtype=t.float32, device=dev),
                                  obs4=t.empty((n, 29), dtype=t.float32, device=dev),
                                  pset=t.empty((n,), dtype=t.uint8, device=dev))

    def _ensure_opponents(self):
        if self._opp is None:
            from .opponents import OpponentPolicies
            dev = self._torch.device("cuda", self.device_index)
            self._opp = OpponentPolicies(self.level, self.args.agent_mode, self._opp_policies_arg, seed=0, device=dev)
            self._ensure_opp_bufs()
        return self._opp

    def opponent_actions(self, actions):
        """First half of a level-4/5 step (hh_step_begin) + the batched opponent networks.
        Returns the opponents' int32 [N,2,4] actions; leaves the env mid-step until hh_step_finish."""
        if self.opponent_action_fn is None:
            opp = self._ensure_opponents()
        else:
            self._ensure_opp_bufs()
        ob = self._opp_bufs
        nat.check(nat.lib().hh_step_begin(self._h, actions.data_ptr(), ob["obs3"].data_ptr(), ob["obs4"].data_ptr(),
                                          ob["pset"].data_ptr(), self._stream()), "hh_step_begin")
        if self.opponent_action_fn is not None:
            self.last_opp_actions = self.opponent_action_fn(ob).contiguous()
        elif self.fused_opponents and ob["obs3"].is_cuda:
            self.last_opp_actions = opp.act_fused(ob["obs3"], ob["obs4"], ob["pset"], self.opponent_precision).contiguous()
        else:
            self.last_opp_actions = opp.act(ob["obs3"], ob["obs4"], ob["pset"]).contiguous()
        return self.last_opp_actions

    # ------------------------------------------------------------------ host (numpy) API
    def reset_host(self, mask: np.ndarray | None = None):
        o1 = np.empty((self.n_arenas, self.obs_dim[0]), np.float32)
        o2 = np.empty((self.n_arenas, self.obs_dim[1]), np.float32)
        mp = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
            mp = mask.ctypes.data
        nat.check(nat.lib().hh_reset_host(self._h, mp, o1.ctypes.data, o2.ctypes.data), "hh_reset_host")
        return o1, o2

    def host_buffers(self):
        """(actions i32 [N,2,4], obs1, obs2, rew f32, done u8) numpy views of the handle's pinned slab: write the
        actions in place and call `step_host(actions, out=(obs1, obs2, rew, done))` for a zero-copy host step."""
        ptrs = [nat.VP() for _ in range(5)]
        nat.check(nat.lib().hh_host_buffers(self._h, *[ctypes.byref(p) for p in ptrs]), "hh_host_buffers")
        n, (d1, d2) = self.n_arenas, self.obs_dim

        def view(p, ctype, shape):
            cnt = int(np.prod(shape))
            return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctype)), shape=(cnt,)).reshape(shape)

        return (view(ptrs[0], ctypes.c_int32, (n, 2, 4)), view(ptrs[1], ctypes.c_float, (n, d1)),
                view(ptrs[2], ctypes.c_float, (n, d2)), view(ptrs[3], ctypes.c_float, (n, 2)),
                view(ptrs[4], ctypes.c_uint8, (n,)))

    def set_host_mode(self, mode: str):
        """'zerocopy' (default: the kernel reads / writes the pinned slab directly), 'staged' (H2D, launch, D2H) or 'pipelined'
        (two half-batch launches in one CUDA graph, the first half's observations travel under the second half's kernel; levels
        1-3 with >= 2 048 arenas, zero-copy otherwise; bit-identical, measured slower than zero-copy at 8 192 arenas)."""
        modes = {"staged": 0, "zerocopy": 1, "pipelined": 2}
        if mode not in modes:
            raise ValueError("mode must be 'staged', 'zerocopy' or 'pipelined'")
        nat.check(nat.lib().hh_set_host_mode(self._h, modes[mode]), "hh_set_host_mode")

    def step_host(self, actions: np.ndarray, out=None):
        """actions: int32 [N, 2, 4] host array -> (obs1, obs2, rew, done) host arrays.
        Host->device and device->host copies happen inside the call."""
        a = np.ascontiguousarray(actions, np.int32)
        if a.size != self.n_arenas * 8:
            raise ValueError("actions must have shape [N, 2, 4]")
        if self.level >= 4:   # the opponent networks run on the device: go through the tensor API
            t = self._ensure_torch() and self._torch
            o1, o2, r, d = self.step(t.from_numpy(a.reshape(self.n_arenas, 2, 4)).cuda(self.device_index))
            return o1.cpu().numpy(), o2.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        if out is None:
            out = (np.empty((self.n_arenas, self.obs_dim[0]), np.float32),
                   np.empty((self.n_arenas, self.obs_dim[1]), np.float32),
                   np.empty((self.n_arenas, 2), np.float32), np.empty((self.n_arenas,), np.uint8))
        o1, o2, r, d = out
        nat.check(nat.lib().hh_step_host(self._h, a.ctypes.data, o1.ctypes.data, o2.ctypes.data, r.ctypes.data,
                                         d.ctypes.data), "hh_step_host")
        return o1, o2, r, d

    def send_actions_host(self, actions: np.ndarray):
        """First half of step_host (RLlib BaseEnv.send_actions): enqueue the step, return immediately."""
        a = np.ascontiguousarray(actions, np.int32)
        if a.size != self.n_arenas * 8:
            raise ValueError("actions must have shape [N, 2, 4]")
        if self.level >= 4:
            raise ValueError("send_actions_host / poll_host: levels 1-3 (levels 4/5 step through the tensor API)")
        nat.check(nat.lib().hh_step_host_begin(self._h, a.ctypes.data), "hh_step_host_begin")

    def poll_host(self, out=None):
        """Second half (RLlib BaseEnv.poll): wait for the enqueued step -> (obs1, obs2, rew, done) host arrays."""
        if out is None:
            out = (np.empty((self.n_arenas, self.obs_dim[0]), np.float32),
                   np.empty((self.n_arenas, self.obs_dim[1]), np.float32),
                   np.empty((self.n_arenas, 2), np.float32), np.empty((self.n_arenas,), np.uint8))
        o1, o2, r, d = out
        nat.check(nat.lib().hh_step_host_end(self._h, o1.ctypes.data, o2.ctypes.data, r.ctypes.data, d.ctypes.data),
                  "hh_step_host_end")
        return o1, o2, r, d

    # ------------------------------------------------------------------ state access
    def _alloc_view(self):
        n = self.n_arenas
        arrays = {name: np.zeros(n * per, dtype=np.dtype(dt)) for name, dt, per in nat.STATE_FIELDS}
        view = nat.HHStateView(**{k: v.ctypes.data for k, v in arrays.items()})
        return arrays, view

    def get_state(self) -> dict:
        arrays, view = self._alloc_view()
        nat.check(nat.lib().hh_get_state(self._h, ctypes.byref(view)), "hh_get_state")
        n = self.n_arenas
        return {name: arrays[name].reshape(n, per) if per > 1 else arrays[name]
                for name, _, per in nat.STATE_FIELDS}

    def set_state(self, state: dict):
        arrays, view = self._alloc_view()
        for name, dt, per in nat.STATE_FIELDS:
            arrays[name][:] = np.asarray(state[name], dtype=np.dtype(dt)).reshape(-1)
        nat.check(nat.lib().hh_set_state(self._h, ctypes.byref(view)), "hh_set_state")

    @property
    def launch_count(self) -> int:
        return int(nat.lib().hh_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            nat.lib().hh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
