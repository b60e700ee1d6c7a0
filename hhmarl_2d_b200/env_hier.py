"""VecHighLevelEnv -- the reference's HighLevelEnv (envs/env_hier.py; 3-vs-3 commander environment, BASELINE
config 5) for N arenas in lock-step on the GPU (csrc/hh_hier.cu through the hh_hier_* C ABI).

A commander step runs _action_assess, then up to 16 low-level sub-steps in which EVERY live aircraft queries a
frozen fight or escape policy (env_hier.py:114-140).  The reference does that with one batch-1 forward per
aircraft per sub-step; here each sub-step is two kernel launches around two batched forwards (agents, then
opponents -- the opponents observe the agents' fresh fire decisions), arenas whose loop has ended idle.

Low-level policies follow the reference's container (env_base.py:332-341): {"fight_1", "fight_2", "escape_1",
"escape_2"} -> networks of hhmarl_2d_b200.models (seeded stand-ins when no weights are given: the reference
loads pickled RLlib models that are not in its repository).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _native as nat
from . import models as M
from .spaces import Box, Discrete

OBS_HL, N_ACTIONS_HL = 34, 3   # env_hier.py:20-37
# keys of the info dict of HHMARLBaseEnv.step under args.eval_info, in the reference's order (env_base.py:104-106)
EVAL_INFO_KEYS = ("agents_win", "opps_win", "draw", "agent_fight", "agent_escape", "opp_fight", "opp_escape",
                  "agent_steps", "opp_steps", "opp1", "opp2", "opp3")


def default_lowlevel_policies(seed: int = 0, device="cpu"):
    import torch  # noqa: F401
    f1, f2 = M.build_policy_pair("fight")
    e1, e2 = M.build_policy_pair("escape")
    for k, m in enumerate((f1, f2, e1, e2)):
        M.fill_from_seed(m, seed + 200 + k)
        m.to(device).eval()
    return {"fight_1": f1, "fight_2": f2, "escape_1": e1, "escape_2": e2}


def make_hier_args(horizon=500, map_size=0.5, rew_scale=1.0, glob_frac=0.0, friendly_kill=True,
                   hier_action_assess=True, hier_opp_fight_ratio=75, level=1, eval_info=False):
    """Config(1) defaults of the reference (config.py:17-57, 98); eval_info=True is Config(2) (config.py:49)."""
    from argparse import Namespace
    return Namespace(level=level, horizon=horizon, agent_mode="fight", num_agents=3, num_opps=3, total_num=6,
                     map_size=map_size, rew_scale=rew_scale, glob_frac=glob_frac, friendly_kill=friendly_kill,
                     hier_action_assess=hier_action_assess, hier_opp_fight_ratio=hier_opp_fight_ratio, eval_hl=True,
                     eval_info=eval_info)


class VecHighLevelEnv:
    def __init__(self, n_arenas: int, args=None, device: int = 0, seed: int = 0, arena_base: int = 0,
                 autoreset: bool = True, lowlevel_policies=None, groups: int | None = None):
        """`groups`: arenas are independent (and an arena's random streams are keyed by arena_base + its index), so the batch
        can live in `groups` native handles of consecutive arenas that step on their own streams inside the commander step's
        graph: one group's env kernels run while another group's policy forward does, and every forward launch stays within
        one round of row tiles (8 192 arenas: 3.67 -> 3.14 ms per commander step with 4 groups, profiles/r3e_hier_groups.txt).  Same
        per-arena results.  None = 4 from 8 192 arenas, 2 from 4 096."""
        import torch
        self._torch = torch
        self.args = args if args is not None else make_hier_args()
        a = self.args
        self.n_arenas, self.device_index = int(n_arenas), int(device)
        if groups is None:
            groups = int(os.environ.get("HH_HIER_GROUPS", "0")) or (4 if self.n_arenas >= 8192 else 2 if self.n_arenas >= 4096 else 1)
        groups = max(1, min(int(groups), self.n_arenas))
        per = (self.n_arenas + groups - 1) // groups
        self._bounds = [(g * per, min(self.n_arenas, (g + 1) * per)) for g in range(groups) if g * per < self.n_arenas]
        self.groups = len(self._bounds)
        self._hs = []
        for lo, hi in self._bounds:
            cfg = nat.HHHierConfig(horizon=a.horizon, level=a.level, friendly_kill=int(bool(a.friendly_kill)),
                                   hier_action_assess=int(bool(a.hier_action_assess)),
                                   hier_opp_fight_ratio=int(a.hier_opp_fight_ratio), autoreset=int(bool(autoreset)),
                                   map_size=float(a.map_size), rew_scale=float(a.rew_scale), glob_frac=float(a.glob_frac),
                                   seed=int(seed), arena_base=int(arena_base) + lo)
            h = nat.VP()
            nat.check(nat.lib().hh_hier_create(ctypes.byref(cfg), hi - lo, self.device_index, ctypes.byref(h)), "hh_hier_create")
            self._hs.append(h)
        self._h = self._hs[0]
        self.observation_space = Box(np.zeros(OBS_HL), np.ones(OBS_HL), dtype=np.float32)   # env_hier.py:36
        self.action_space = Discrete(N_ACTIONS_HL)                                            # env_hier.py:37
        self._agent_ids = {1, 2, 3}
        dev = torch.device("cuda", self.device_index)
        self.dev = dev
        n = self.n_arenas
        self.policies = lowlevel_policies if lowlevel_policies is not None else default_lowlevel_policies(0, dev)
        for m in self.policies.values():       # supplied policies (checkpoint.load_highlevel_policies defaults to the CPU) move to
            m.to(dev).eval()                   # the env's device: the fused forward hands their packed weights to the kernel
        self.obs = torch.empty((n, 3, OBS_HL), dtype=torch.float32, device=dev)
        self.rew = torch.empty((n, 3), dtype=torch.float32, device=dev)
        self.done = torch.empty((n,), dtype=torch.uint8, device=dev)
        self.substeps = torch.empty((n,), dtype=torch.int32, device=dev)
        self.ll_obs = torch.zeros((n, 6, 30), dtype=torch.float32, device=dev)
        self.ll_info = torch.zeros((n, 6), dtype=torch.uint8, device=dev)
        self.ll_act = torch.zeros((n, 6, 4), dtype=torch.int32, device=dev)
        # args.eval_info (config.py:49, env_base.py:91-107): when set, step() also fills `self.info` (int32 [N,12],
        # columns EVAL_INFO_KEYS) with the reference's per-step info dict; hhmarl_2d_b200.evaluation sums it
        self.eval_info = bool(getattr(self.args, "eval_info", False))
        self.info = torch.zeros((n, len(EVAL_INFO_KEYS)), dtype=torch.int32, device=dev)
        self.tick_hook = None   # callable(sub_step) run after every hh_hier_tick (trace.HierTraceRecorder); None = no cost
        self.trace = None   # set to a list to record (phase, ll_obs, ll_info, ll_act) per sub-step (tests)
        self.fused_policies = True   # the frozen low-level actors through csrc/hh_policy.cu (False: torch forward)
        self.policy_precision = 2    # 2: tcgen05 path (fp32-equivalent logits before the argmax), 0: 3xTF32 on mma.sync, 1: plain TF32
        self._fused = None
        # use_cuda_graph: a commander step is ~70 launches with no host synchronisation; after `graph_warmup` eager steps it is
        # captured once and replayed (commander actions pass through a buffer that stays).  Eager when a trace / tick hook /
        # the torch policies are in use.
        self.use_cuda_graph = os.environ.get("HH_HIER_GRAPH", "1") != "0"
        self.graph_warmup = 2
        self._graph, self._eager_steps, self._ca_static = None, 0, None

    def _stream(self):
        return self._torch.cuda.current_stream(self.device_index).cuda_stream

    def reset(self, mask=None):
        for h, (lo, hi) in zip(self._hs, self._bounds):
            mp = None if mask is None else mask[lo:hi].data_ptr()
            nat.check(nat.lib().hh_hier_reset(h, mp, self.obs[lo:hi].data_ptr(), self._stream()), "hh_hier_reset")
        return self.obs

    # frozen low-level policies, batched: env_base.py:349-398 (per-head argmax of the actor).  Which network a
    # unit uses (fight/escape x aircraft type) is fixed for the whole commander step, so the row lists are built
    # once per step (on the device: hh_hier_policy_rows) and reused by all 16 sub-steps; rows of units that do not query right now
    # (dead, or arena already out of its sub-step loop) are computed and ignored by the kernels.
    _KINDS = (("fight", 1, 0), ("fight", 2, 4), ("escape", 1, 2), ("escape", 2, 6))
    _DIMS = {("fight", 1): 26, ("fight", 2): 24, ("escape", 1): 30, ("escape", 2): 29}

    def _build_rows(self, g: int = 0):
        t = self._torch
        if self.fused_policies:
            # device-built lists (hh_hier_policy_rows), per group of arenas: no host synchronisation anywhere in the commander step
            if getattr(self, "_rows_dev", None) is None:
                self._rows_dev = [t.zeros((8, (hi - lo) * 3), dtype=t.int32, device=self.dev) for lo, hi in self._bounds]
                self._ranges_dev = [t.zeros((8, 2), dtype=t.int32, device=self.dev) for _ in self._bounds]
            lo, hi = self._bounds[g]
            nat.check(nat.lib().hh_hier_policy_rows(self._hs[g], self.ll_info[lo:hi].data_ptr(), self._rows_dev[g].data_ptr(),
                                                    self._ranges_dev[g].data_ptr(), self._stream()), "hh_hier_policy_rows")
            self._rows = [(mode, ac, first, 2 * k + (first // 3)) for k, (mode, ac, _) in enumerate(self._KINDS) for first in (0, 3)]
            return
        kind = (self.ll_info & 6).reshape(-1)
        live = (self.ll_info & 8).reshape(-1) != 0     # alive at the start of this commander step (hh_hier_begin)
        self._rows = []
        for mode, ac, bits in self._KINDS:
            for first in (0, 3):
                unit = t.arange(self.n_arenas * 6, device=self.dev) % 6
                sel = (kind == bits) & live & (unit >= first) & (unit < first + 3)
                idx = t.nonzero(sel, as_tuple=False).flatten()
                self._rows.append((mode, ac, first, idx.to(t.int32) if self.fused_policies else idx))

    def _policy_key(self, mode: str, ac: int, first: int) -> str:
        """env_base.py:385-390: every aircraft uses "{mode}_{ac_type}"; only in the low-level evaluation mode
        (args.eval_hl False, env_base.py:343-346) the opponents (units 4-6) fight with "fight_{ac_type}_opp"."""
        if first == 3 and mode == "fight" and not getattr(self.args, "eval_hl", True):
            return f"fight_{ac}_opp"
        return f"{mode}_{ac}"

    def _infer_fused(self, first: int, g: int = 0):
        """All (mode x aircraft type) row lists of this half-step (of group g's arenas) as chains of ONE hh_policy_forward_ex
        launch: gather by row index, actor forward on the tensor cores, per-head argmax written straight into ll_act."""
        from .fused_forward import FusedActor, run_chains
        if self._fused is None:
            self._fused = {}
        lo, hi = self._bounds[g]
        obs_flat, act_flat = self.ll_obs[lo:hi].reshape(-1, 30), self.ll_act[lo:hi].reshape(-1, 4)
        fills = []
        cap = (hi - lo) * 3
        for mode, ac, f, lst in self._rows:
            if f != first:
                continue
            key = self._policy_key(mode, ac, first)
            if key not in self._fused:
                self._fused[key] = FusedActor(self.policies[key])
            fa = self._fused[key]
            fills.append(lambda c, fa=fa, lst=lst: fa.fill_chain(c, obs_flat, cap, act_out=act_flat, rows=self._rows_dev[g],
                                                                 range_dev=self._ranges_dev[g][lst]))
        run_chains(fills, self.dev, self.policy_precision)

    def _infer(self, first: int, g: int = 0):
        if self.fused_policies:
            return self._infer_fused(first, g)
        t = self._torch
        obs_flat, act_flat = self.ll_obs.reshape(-1, 30), self.ll_act.reshape(-1, 4)
        with t.no_grad():
            for mode, ac, f, idx in self._rows:
                if f != first or idx.numel() == 0:
                    continue
                x = obs_flat.index_select(0, idx)[:, :self._DIMS[(mode, ac)]]
                a = M.deterministic_actions(self.policies[self._policy_key(mode, ac, first)].actor(x), ac).to(t.int32)
                if a.shape[1] == 3:
                    a = t.nn.functional.pad(a, (0, 1))
                act_flat.index_copy_(0, idx, a)

    def step(self, commander_actions):
        """commander_actions: int32 CUDA tensor [N, 3] in {0: escape, 1: nearest opponent, 2: second nearest}.
        Returns (obs [N,3,34], rew [N,3], done [N] u8); `self.substeps` holds the sub-step counts."""
        t = self._torch
        assert commander_actions.is_cuda and commander_actions.dtype == t.int32 and commander_actions.numel() == self.n_arenas * 3
        if not (self.use_cuda_graph and self.fused_policies and self.trace is None and self.tick_hook is None):
            return self._step(commander_actions)
        if self._graph is None and self._eager_steps < self.graph_warmup:
            self._eager_steps += 1
            return self._step(commander_actions)
        if self._ca_static is None:
            self._ca_static = t.empty((self.n_arenas, 3), dtype=t.int32, device=self.dev)
        self._ca_static.copy_(commander_actions.reshape(self.n_arenas, 3))
        if self._graph is None:
            t.cuda.synchronize(self.dev)
            self._graph = t.cuda.CUDAGraph()
            with t.cuda.graph(self._graph):           # records, does not run
                self._step(self._ca_static)
        self._graph.replay()
        return self.obs, self.rew, self.done

    def _step(self, commander_actions):
        """One commander step.  Under a graph capture (no trace / hook, fused policies) every group of arenas runs its own chain of
        launches on its own stream (group-major); otherwise the groups take every phase one after the other on the current
        stream (phase-major: a trace or tick hook then sees all arenas at the same phase)."""
        t = self._torch
        ca = commander_actions.contiguous().reshape(self.n_arenas, 3)
        serial = not (self.groups > 1 and self.fused_policies and self.trace is None and self.tick_hook is None
                      and t.cuda.is_current_stream_capturing())
        if not serial:
            cur = t.cuda.current_stream(self.dev)
            if getattr(self, "_gstreams", None) is None:
                self._gstreams = [t.cuda.Stream(self.dev) for _ in self._bounds]
            for g, gs in enumerate(self._gstreams):
                gs.wait_stream(cur)
                with t.cuda.stream(gs):
                    self._begin(g, ca)
                    for s in range(16):
                        self._infer(0, g)
                        self._agents(g)
                        self._infer(3, g)
                        self._tick(g)
                    self._end(g)
            for gs in self._gstreams:
                cur.wait_stream(gs)
            return self.obs, self.rew, self.done
        G = range(self.groups)
        for g in G:
            self._begin(g, ca)
        if not self.fused_policies:
            self._build_rows()
        for s in range(16):   # n_sub_steps = 15 -> at most 16 iterations (env_hier.py:33,125)
            for g in (G if self.fused_policies else (0,)):
                self._infer(0, g)
            if self.trace is not None:
                self.trace.append(("agents", self.ll_obs.cpu().numpy().copy(), self.ll_info.cpu().numpy().copy(), None))
            for g in G:
                self._agents(g)
            for g in (G if self.fused_policies else (0,)):
                self._infer(3, g)
            if self.trace is not None:
                self.trace.append(("opps", self.ll_obs.cpu().numpy().copy(), self.ll_info.cpu().numpy().copy(),
                                   self.ll_act.cpu().numpy().copy()))
            for g in G:
                self._tick(g)
            if self.tick_hook is not None:
                self.tick_hook(s)
        for g in G:
            self._end(g)
        return self.obs, self.rew, self.done

    # the native calls of one group of arenas (its handle, its rows of the shared tensors)
    def _begin(self, g, ca):
        lo, hi = self._bounds[g]
        nat.check(nat.lib().hh_hier_begin(self._hs[g], ca[lo:hi].data_ptr(), self.ll_obs[lo:hi].data_ptr(),
                                          self.ll_info[lo:hi].data_ptr(), self._stream()), "hh_hier_begin")
        if self.fused_policies:
            self._build_rows(g)

    def _agents(self, g):
        lo, hi = self._bounds[g]
        nat.check(nat.lib().hh_hier_agents(self._hs[g], self.ll_act[lo:hi].data_ptr(), self.ll_obs[lo:hi].data_ptr(),
                                           self.ll_info[lo:hi].data_ptr(), self._stream()), "hh_hier_agents")

    def _tick(self, g):
        lo, hi = self._bounds[g]
        nat.check(nat.lib().hh_hier_tick(self._hs[g], self.ll_act[lo:hi].data_ptr(), self.ll_obs[lo:hi].data_ptr(),
                                         self.ll_info[lo:hi].data_ptr(), self._stream()), "hh_hier_tick")

    def _end(self, g):
        lo, hi = self._bounds[g]
        L, h, st = nat.lib(), self._hs[g], self._stream()
        if self.eval_info:   # before hh_hier_end: its auto-reset replaces a finished episode
            nat.check(L.hh_hier_eval_info(h, self.info[lo:hi].data_ptr(), st), "hh_hier_eval_info")
        nat.check(L.hh_hier_end(h, self.obs[lo:hi].data_ptr(), self.rew[lo:hi].data_ptr(), self.done[lo:hi].data_ptr(),
                                self.substeps[lo:hi].data_ptr(), st), "hh_hier_end")

    def get_state(self):
        arr = (nat.HHHierArena * self.n_arenas)()
        sz = ctypes.sizeof(nat.HHHierArena)
        for h, (lo, hi) in zip(self._hs, self._bounds):
            nat.check(nat.lib().hh_hier_get_state(h, ctypes.cast(ctypes.addressof(arr) + lo * sz, nat.VP)), "hh_hier_get_state")
        return arr

    @property
    def launch_count(self):
        return sum(int(nat.lib().hh_hier_launch_count(h)) for h in self._hs)

    def close(self):
        for h in getattr(self, "_hs", None) or []:
            nat.lib().hh_hier_destroy(h)
        self._hs, self._h = [], None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CommanderSampler:
    """On-device rollout of the commander policy (train_hier.py): one CommanderGru shared by the three agents
    (policy_mapping_fn -> "commander_policy"), central observation = sorted-key flattening of
    central_critic_observer (train_hier.py:134-165): [act_1_own, act_2, act_3 | obs_1_own | obs_2 | obs_3] (105),
    actions zero while sampling and written back as act / N_OPP_HL afterwards (train_hier.py:131-132);
    GRU state carried per (arena, agent) and zeroed when the arena's episode ends."""

    def __init__(self, env: VecHighLevelEnv, model, fragment_len: int = 8, gamma: float = 0.99, lam: float = 1.0):
        import torch
        self.env, self.model, self.T, self.gamma, self.lam = env, model, fragment_len, gamma, lam
        n, dev = env.n_arenas, env.dev
        f32 = dict(dtype=torch.float32, device=dev)
        self.buf = dict(flat=torch.zeros((fragment_len, n, 3, 105), **f32), actions=torch.zeros((fragment_len, n, 3), dtype=torch.int32, device=dev),
                        logp=torch.zeros((fragment_len, n, 3), **f32), vf=torch.zeros((fragment_len, n, 3), **f32),
                        logits=torch.zeros((fragment_len, n, 3, 3), **f32), rew=torch.zeros((fragment_len, n, 3), **f32),
                        done=torch.zeros((fragment_len, n), dtype=torch.uint8, device=dev),
                        h0=torch.zeros((fragment_len, n, 3, 200), **f32), h1=torch.zeros((fragment_len, n, 3, 200), **f32),
                        adv=torch.zeros((fragment_len, n, 3), **f32), vtarg=torch.zeros((fragment_len, n, 3), **f32),
                        last_vf=torch.zeros((n, 3), **f32), substeps=torch.zeros((fragment_len, n), dtype=torch.int32, device=dev))
        self.h = [torch.zeros((n * 3, 200), **f32), torch.zeros((n * 3, 200), **f32)]
        self.others = torch.tensor([[1, 2], [0, 2], [0, 1]], device=dev)
        self.obs = None

    def _flat(self, obs):   # obs [N,3,34] -> [N,3,105]
        t = self.env._torch
        o = obs[:, self.others]                                      # [N,3,2,34]: the two team-mates in id order
        return t.cat((t.zeros((obs.shape[0], 3, 3), device=obs.device), obs, o[:, :, 0], o[:, :, 1]), dim=2)

    def _forward(self, flat):
        x = flat.reshape(-1, 105)
        d = {"act_1_own": x[:, 0:1], "act_2": x[:, 1:2], "act_3": x[:, 2:3], "obs_1_own": x[:, 3:37], "obs_2": x[:, 37:71],
             "obs_3": x[:, 71:105]}
        logits, new_h = self.model({"obs": d}, self.h, self.env._torch.ones(x.shape[0], dtype=self.env._torch.int32))
        return logits, self.model.value_function(), new_h

    def collect(self):
        t = self.env._torch
        b = self.buf
        n = self.env.n_arenas
        with t.no_grad():
            if self.obs is None:
                self.obs = self.env.reset().clone()
            for k in range(self.T):
                flat = self._flat(self.obs)
                b["h0"][k], b["h1"][k] = self.h[0].view(n, 3, 200), self.h[1].view(n, 3, 200)
                logits, vf, new_h = self._forward(flat)
                g = -t.log(-t.log(t.rand_like(logits).clamp_(1e-20, 1 - 1e-7)))
                a = t.argmax(logits + g, dim=-1)
                lsm = t.log_softmax(logits, dim=-1)
                b["flat"][k], b["logits"][k], b["vf"][k] = flat, logits.view(n, 3, 3), vf.view(n, 3)
                b["actions"][k] = a.view(n, 3).to(t.int32)
                b["logp"][k] = lsm.gather(1, a[:, None])[:, 0].view(n, 3)
                obs, rew, done = self.env.step(b["actions"][k].contiguous())
                b["rew"][k], b["done"][k], b["substeps"][k] = rew, done, self.env.substeps
                keep = (1 - done.float()).repeat_interleave(3)[:, None]
                self.h = [new_h[0] * keep, new_h[1] * keep]          # new episode -> initial state (zeros)
                self.obs = obs.clone()
            _, last_v, _ = self._forward(self._flat(self.obs))
            b["last_vf"].copy_(last_v.view(n, 3))
            st = t.cuda.current_stream(self.env.dev).cuda_stream
            nat.check(nat.lib().hh_gae_agents(self.T, n, 3, b["rew"].data_ptr(), b["vf"].data_ptr(), b["last_vf"].data_ptr(),
                                              b["done"].data_ptr(), self.gamma, self.lam, b["adv"].data_ptr(),
                                              b["vtarg"].data_ptr(), st), "hh_gae_agents")
            act = b["actions"].float() / 2.0                          # N_OPP_HL = 2 (train_hier.py:131-132)
            b["flat"][..., 0] = act
            b["flat"][..., 1] = act[:, :, self.others[:, 0]]
            b["flat"][..., 2] = act[:, :, self.others[:, 1]]
        return b
