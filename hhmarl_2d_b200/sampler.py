"""On-device sampler: policy forward -> MultiCategorical sampling -> fused env step, for all arenas in
lock-step, without touching the host.  Replaces RLlib's RolloutWorker loop:

  central_critic_observer (train_hetero.py:162-181)  -> flat [act_1_own | act_2 | obs_1_own | obs_2], actions 0
  Policy.compute_actions -> Fight1/Fight2.forward     -> logits, vf_preds (critic sees ZERO actions here, A.6.15)
  env.step(action_dict)                              -> VecLowLevelEnv.step (one kernel launch)
  CustomCallback.on_postprocess_trajectory (:120-160) -> real, scaled actions written into the first 7 columns
  RLlib GAE postprocessing                           -> hh_gae kernel
"""
from __future__ import annotations

import os

import torch

from . import _native as nat
from . import models as M

ACT_SCALE = torch.tensor([12.0, 8.0, 1.0, 1.0])  # train_hetero.py:143-146: a0/12, a1/8, a2, a3


def multicategorical_sample(logits: torch.Tensor, splits, explore: bool = True):
    """Per-head categorical sample (Gumbel-max) or argmax; returns (actions int64 [B,H], logp [B])."""
    acts, logp, o = [], 0.0, 0
    for n in splits:
        seg = logits[:, o:o + n]
        lsm = torch.log_softmax(seg, dim=-1)
        if explore:
            g = -torch.log(-torch.log(torch.rand_like(seg).clamp_(1e-20, 1.0 - 1e-7)))
            a = torch.argmax(seg + g, dim=-1)
        else:
            a = torch.argmax(seg, dim=-1)
        acts.append(a)
        logp = logp + lsm.gather(1, a[:, None])[:, 0]
        o += n
    return torch.stack(acts, dim=1), logp


class _FusedMultiCat(torch.autograd.Function):
    """The three MultiCategorical terms of one policy as ONE forward and ONE backward kernel (hh_multicat_forward / _backward)
    instead of ~35 element-wise torch kernels per head and direction.  CUDA float32 only."""

    @staticmethod
    def forward(ctx, logits, old_logits, actions, splits):
        import ctypes
        n = logits.shape[0]
        logits = logits.contiguous()
        w = (ctypes.c_int32 * len(splits))(*[int(v) for v in splits])
        out = torch.empty((3, n), dtype=torch.float32, device=logits.device)
        st = torch.cuda.current_stream(logits.device).cuda_stream
        nat.check(nat.lib().hh_multicat_forward(n, len(splits), w, logits.data_ptr(), logits.stride(0), old_logits.data_ptr(),
                                                old_logits.stride(0), actions.data_ptr(), actions.stride(0), out[0].data_ptr(),
                                                out[1].data_ptr(), out[2].data_ptr(), st), "hh_multicat_forward")
        ctx.save_for_backward(logits, old_logits, actions)
        ctx.splits = tuple(int(v) for v in splits)
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_logp, g_ent, g_kl):
        import ctypes
        logits, old_logits, actions = ctx.saved_tensors
        n = logits.shape[0]
        w = (ctypes.c_int32 * len(ctx.splits))(*ctx.splits)
        g = torch.stack([g_logp, g_ent, g_kl]).contiguous()     # (also materialises expanded / zero gradients)
        gz = torch.empty((n, sum(ctx.splits)), dtype=torch.float32, device=logits.device)
        st = torch.cuda.current_stream(logits.device).cuda_stream
        nat.check(nat.lib().hh_multicat_backward(n, len(ctx.splits), w, logits.data_ptr(), logits.stride(0), old_logits.data_ptr(),
                                                 old_logits.stride(0), actions.data_ptr(), actions.stride(0), g[0].data_ptr(),
                                                 g[1].data_ptr(), g[2].data_ptr(), gz.data_ptr(), st), "hh_multicat_backward")
        return gz, None, None, None


def multicategorical_logp_entropy_kl(logits, actions, splits, old_logits=None):
    """log-prob of `actions`, entropy and (optionally) KL(old || new), summed over the heads."""
    if (logits.is_cuda and old_logits is not None and logits.dtype == torch.float32 and old_logits.dtype == torch.float32
            and actions.dtype == torch.int32 and old_logits.stride(-1) == 1 and actions.stride(-1) == 1
            and logits.shape[1] == sum(splits) and len(splits) <= 4):
        return _FusedMultiCat.apply(logits, old_logits, actions, tuple(splits))
    logp = ent = kl = 0.0
    o = 0
    for h, n in enumerate(splits):
        lsm = torch.log_softmax(logits[:, o:o + n], dim=-1)
        p = lsm.exp()
        logp = logp + lsm.gather(1, actions[:, h:h + 1].long())[:, 0]
        ent = ent - (p * lsm).sum(-1)
        if old_logits is not None:
            lo = torch.log_softmax(old_logits[:, o:o + n], dim=-1)
            kl = kl + (lo.exp() * (lo - lsm)).sum(-1)
        o += n
    return logp, ent, kl


class TorchPolicy:
    """Minimal mirror of RLlib's Policy for one aircraft type: `compute_actions(obs_batch, ...)` takes the
    flattened central observation and returns (actions, state_out, extra_fetches) with the RLlib keys."""

    def __init__(self, model: torch.nn.Module, ac_type: int):
        self.model, self.ac_type = model, ac_type
        self.splits = M.ACTION_SPLITS[ac_type]
        self._fused = None           # (FusedPolicyPair, index): set by VecSampler / attach_fused

    def attach_fused(self, pair, index: int):
        """Route compute_actions through the fused tcgen05 forward of `pair` (fused_forward.FusedPolicyPair, precision 2)
        instead of the eager torch modules; `index` 0 = ac1_policy, 1 = ac2_policy."""
        self._fused = (pair, index)

    @torch.no_grad()
    def compute_actions(self, obs_batch, state_batches=None, prev_action_batch=None, prev_reward_batch=None,
                        explore=True, **kw):
        """RLlib's Policy.compute_actions contract (train_hetero.py:200-205, 242): obs_batch = the flattened central
        observation [B, 57|66] -> (actions [B, heads], state_out = [], {action_logp, action_dist_inputs, vf_preds})."""
        if self._fused is not None and obs_batch.is_cuda:
            pair, idx = self._fused
            logits, vf = pair.forward_one(idx, obs_batch.contiguous())
            logits, vf = logits.clone(), vf.clone()
        else:
            logits, vf = self.model.forward_flat(obs_batch)
        actions, logp = multicategorical_sample(logits, self.splits, explore)
        return actions, [], {"action_logp": logp, "action_dist_inputs": logits, "vf_preds": vf}

    def compute_single_action(self, obs, explore=False, **kw):
        a, s, x = self.compute_actions(obs[None], explore=explore)
        return a[0], s, {k: v[0] for k, v in x.items()}


class VecSampler:
    """Collects rollout fragments of T lock-step ticks from a VecLowLevelEnv with two policies
    (ac1_policy for agent 1, ac2_policy for agent 2; policy_mapping_fn of train_hetero.py:240)."""

    def __init__(self, env, policy1: TorchPolicy, policy2: TorchPolicy, fragment_len: int = 64,
                 gamma: float = 0.99, lam: float = 0.95, use_cuda_graph: bool = True, packed: bool = True,
                 allow_tf32: bool = False, fused: str | None = "tc", groups: int | None = None, n_buffers: int = 1):
        """`fused`: "tc" (default: csrc/hh_policy_tc.cu, tcgen05 / TMEM forward, fp32-equivalent, one launch per tick),
        "3xtf32" / "tf32" (csrc/hh_policy.cu on mma.sync: fp32-equivalent / plain TF32 products) or None (cuBLAS:
        `packed` / per-layer torch forward).
        `groups`: arenas are independent, so the batch can be advanced as `groups` equal parts on their own streams inside the
        fragment's graph -- one part's env step and sampling glue run while the other parts' policy forwards do, and the
        forward's rounds of row tiles interleave instead of ending with a partly filled round (8 192 arenas: 0.112 -> 0.100 ms
        per tick with 8 groups, profiles/r2zk_sampler_groups.txt).  Same per-arena results (tested bit for bit).
        None = 8 from 8 192 arenas, 4 from 4 096 (fused forward, levels 1-3, arenas a multiple of 128 x groups), else 1
        (HH_SAMPLER_GROUPS overrides the default).
        `n_buffers`: 2 = two sets of rollout buffers used in turn (one CUDA graph each), so that collect_host() can copy
        fragment k to the host while fragment k + 1 is sampled."""
        self.env, self.p1, self.p2, self.T = env, policy1, policy2, fragment_len
        self.allow_tf32 = allow_tf32
        self.packed = None
        if fused is not None:
            if fused not in ("tc", "3xtf32", "tf32"):
                raise ValueError("fused must be 'tc', '3xtf32', 'tf32' or None")
            from .fused_forward import FusedPolicyPair
            self.packed = FusedPolicyPair(policy1.model, policy2.model, precision={"tc": 2, "3xtf32": 0, "tf32": 1}[fused])
            if fused == "tc":          # the policies' own compute_actions run the same kernel
                policy1.attach_fused(self.packed, 0)
                policy2.attach_fused(self.packed, 1)
        elif packed:
            from .fused_forward import PackedPolicyPair
            self.packed = PackedPolicyPair(policy1.model, policy2.model)
        self.gamma, self.lam = gamma, lam
        n, T = env.n_arenas, fragment_len
        d1, d2 = env.obs_dim
        dev = torch.device("cuda", env.device_index)
        self.dev = dev
        f32 = dict(dtype=torch.float32, device=dev)
        if n_buffers not in (1, 2):
            raise ValueError("n_buffers must be 1 or 2")
        self.bufs = [dict(
            flat1=torch.zeros((T, n, 7 + d1 + d2), **f32), flat2=torch.zeros((T, n, 7 + d1 + d2), **f32),
            actions=torch.zeros((T, n, 2, 4), dtype=torch.int32, device=dev),
            logp=torch.zeros((T, n, 2), **f32), vf=torch.zeros((T, n, 2), **f32),
            logits1=torch.zeros((T, n, sum(policy1.splits)), **f32),
            logits2=torch.zeros((T, n, sum(policy2.splits)), **f32),
            rew=torch.zeros((T, n, 2), **f32), done=torch.zeros((T, n), dtype=torch.uint8, device=dev),
            adv=torch.zeros((T, n, 2), **f32), vtarg=torch.zeros((T, n, 2), **f32),
            last_vf=torch.zeros((n, 2), **f32)) for _ in range(n_buffers)]
        self.buf = self.bufs[0]            # the set the next / last fragment uses
        self._bi = 0
        self._host, self._copy_stream, self._copy_done = None, None, [None] * n_buffers
        self.cur1 = torch.zeros((n, 7 + d1 + d2), **f32)   # [act_1_own(4) | act_2(3) | obs_1_own | obs_2], actions 0
        self.cur2 = torch.zeros((n, 7 + d1 + d2), **f32)   # [act_1_own(3) | act_2(4) | obs_1_own | obs_2]
        self.d1, self.d2 = d1, d2
        # hh_sample_actions depends on the LOGITS layout (26 / 24 in both agent modes), hh_pack_central takes d1, d2:
        # the native glue serves fight (26 / 24) and escape (30 / 29) observations alike
        self.native_glue = True
        self.direct = self.native_glue and fused is not None   # kernels write into the buffers themselves (all levels)
        can_group = self.direct and int(env.level) <= 3
        if groups is None:
            groups = int(os.environ.get("HH_SAMPLER_GROUPS", "0")) or (8 if n >= 8192 else 4 if n >= 4096 else 1)
            while groups > 1 and (not can_group or n % (128 * groups) != 0):
                groups //= 2
        if groups < 1 or (groups > 1 and (not can_group or n % (128 * groups) != 0)):
            raise ValueError("groups > 1 needs the fused forward, a level 1-3 env and a multiple of 128 * groups arenas")
        self.groups = groups
        self._bounds = [(g * (n // groups), (g + 1) * (n // groups)) for g in range(groups)]
        self._gstreams = [torch.cuda.Stream(dev) for _ in range(groups)] if groups > 1 else []
        self.ctr = torch.zeros((n, 2), dtype=torch.int32, device=dev)
        self.seed = int(getattr(env, "_cfg").seed) + 0x5A17
        self.scale = ACT_SCALE.to(dev)
        self.use_graph = use_cuda_graph
        self._graphs = [None] * n_buffers
        self._started = False

    # central_critic_observer, train_hetero.py:162-181
    def _set_obs(self, obs1, obs2):
        d1, d2 = self.d1, self.d2
        if self.native_glue:
            st = torch.cuda.current_stream(self.dev).cuda_stream
            nat.check(nat.lib().hh_pack_central(self.env.n_arenas, d1, d2, obs1.data_ptr(), obs2.data_ptr(),
                                                self.cur1.data_ptr(), self.cur2.data_ptr(), st), "hh_pack_central")
            return
        self.cur1[:, 7:7 + d1] = obs1
        self.cur1[:, 7 + d1:] = obs2
        self.cur2[:, 7:7 + d2] = obs2
        self.cur2[:, 7 + d2:] = obs1

    def refresh_policy(self):
        """Call after the learner changed the weights (re-packs in place; a captured graph stays valid)."""
        if self.packed is not None:
            self.packed.refresh()

    def _forward_both(self, f1, f2):
        if self.packed is not None:
            return self.packed.forward(f1, f2)
        l1, v1 = self.p1.model.forward_flat(f1)
        l2, v2 = self.p2.model.forward_flat(f2)
        return l1, v1, l2, v2

    def _tick_direct(self, t):
        """Tick with every kernel writing straight into the rollout buffers (fused forward + native glue): four
        launches per tick -- forward, sample, env step, central-observation packing -- and no copies."""
        b, n, T = self.buf, self.env.n_arenas, self.T
        st = torch.cuda.current_stream(self.dev).cuda_stream
        vf = b["vf"][t]
        self.packed.forward(b["flat1"][t], b["flat2"][t], out=(b["logits1"][t], vf[:, 0], b["logits2"][t], vf[:, 1]))
        act = b["actions"][t]
        nat.check(nat.lib().hh_sample_actions(n, b["logits1"][t].data_ptr(), b["logits2"][t].data_ptr(), self.seed,
                                              int(self.env._cfg.arena_base), self.ctr.data_ptr(), 1, act.data_ptr(),
                                              b["logp"][t].data_ptr(), st), "hh_sample_actions")
        eb = self.env._ensure_torch()
        obs1, obs2, _, _ = self.env.step(act, out=dict(obs1=eb["obs1"], obs2=eb["obs2"], rew=b["rew"][t], done=b["done"][t]))
        nxt1, nxt2 = (b["flat1"][t + 1], b["flat2"][t + 1]) if t + 1 < T else (self.cur1, self.cur2)
        nat.check(nat.lib().hh_pack_central(n, self.d1, self.d2, obs1.data_ptr(), obs2.data_ptr(), nxt1.data_ptr(),
                                            nxt2.data_ptr(), st), "hh_pack_central")

    def _tick_group(self, t, lo, hi):
        """_tick_direct for arenas [lo, hi) on the current stream (kernels of the whole batch, pointed at the range's rows)."""
        b, T, L = self.buf, self.T, nat.lib()
        st = torch.cuda.current_stream(self.dev).cuda_stream
        vf = b["vf"][t]
        lg1, lg2 = b["logits1"][t], b["logits2"][t]
        self.packed.forward(b["flat1"][t][lo:hi], b["flat2"][t][lo:hi], out=(lg1[lo:hi], vf[lo:hi, 0], lg2[lo:hi], vf[lo:hi, 1]))
        act = b["actions"][t]
        nat.check(L.hh_sample_actions(hi - lo, lg1[lo:hi].data_ptr(), lg2[lo:hi].data_ptr(), self.seed,
                                      int(self.env._cfg.arena_base) + lo, self.ctr[lo:hi].data_ptr(), 1, act[lo:hi].data_ptr(),
                                      b["logp"][t][lo:hi].data_ptr(), st), "hh_sample_actions")
        eb = self.env._ensure_torch()
        nxt1, nxt2 = (b["flat1"][t + 1], b["flat2"][t + 1]) if t + 1 < T else (self.cur1, self.cur2)
        # the step kernel writes the next tick's central observation rows itself (hh_step_range_central): three launches per tick
        self.env.step_range(lo, hi - lo, act, out=dict(obs1=eb["obs1"], obs2=eb["obs2"], rew=b["rew"][t], done=b["done"][t]),
                            central=(nxt1, nxt2))

    def _tick(self, t):
        b = self.buf
        if self.direct:
            if int(self.env.level) <= 3:          # one group = the whole batch: forward, sample, step (+ central rows)
                return self._tick_group(t, 0, self.env.n_arenas)
            return self._tick_direct(t)
        l1, v1, l2, v2 = self._forward_both(self.cur1, self.cur2)
        if self.native_glue:
            b["flat1"][t] = self.cur1
            b["flat2"][t] = self.cur2
            b["logits1"][t], b["logits2"][t] = l1, l2
            b["vf"][t, :, 0], b["vf"][t, :, 1] = v1, v2
            act = b["actions"][t]
            st = torch.cuda.current_stream(self.dev).cuda_stream
            nat.check(nat.lib().hh_sample_actions(self.env.n_arenas, b["logits1"][t].data_ptr(), b["logits2"][t].data_ptr(),
                                                  self.seed, int(self.env._cfg.arena_base), self.ctr.data_ptr(), 1,
                                                  act.data_ptr(), b["logp"][t].data_ptr(), st), "hh_sample_actions")
            obs1, obs2, rew, done = self.env.step(act)
            b["rew"][t] = rew
            b["done"][t] = done
            self._set_obs(obs1, obs2)
            return
        a1, lp1 = multicategorical_sample(l1, self.p1.splits, True)
        a2, lp2 = multicategorical_sample(l2, self.p2.splits, True)
        x1 = {"action_logp": lp1, "action_dist_inputs": l1, "vf_preds": v1}
        x2 = {"action_logp": lp2, "action_dist_inputs": l2, "vf_preds": v2}
        b["flat1"][t] = self.cur1
        b["flat2"][t] = self.cur2
        act = b["actions"][t]
        act[:, 0, :] = a1.to(torch.int32)
        act[:, 1, :3] = a2.to(torch.int32)
        b["logp"][t, :, 0], b["logp"][t, :, 1] = x1["action_logp"], x2["action_logp"]
        b["vf"][t, :, 0], b["vf"][t, :, 1] = x1["vf_preds"], x2["vf_preds"]
        b["logits1"][t], b["logits2"][t] = x1["action_dist_inputs"], x2["action_dist_inputs"]
        obs1, obs2, rew, done = self.env.step(act)
        b["rew"][t] = rew
        b["done"][t] = done
        self._set_obs(obs1, obs2)

    def _fragment(self):
        b = self.buf
        st = torch.cuda.current_stream(self.dev).cuda_stream
        D = 7 + self.d1 + self.d2
        if self.direct:
            # the critic sees ZERO actions while sampling: clear what the previous fragment's write-back
            # left in the action columns, then seed tick 0 with the current central observation (one launch)
            nat.check(nat.lib().hh_fragment_prepare(self.T, self.env.n_arenas, D, b["flat1"].data_ptr(), b["flat2"].data_ptr(),
                                                    self.cur1.data_ptr(), self.cur2.data_ptr(), st), "hh_fragment_prepare")
        if self.groups > 1:      # every half runs its own chain of T ticks on its own stream (fork / join around the loop)
            cur = torch.cuda.current_stream(self.dev)
            for (lo, hi), gs in zip(self._bounds, self._gstreams):
                gs.wait_stream(cur)
                with torch.cuda.stream(gs):
                    for t in range(self.T):
                        self._tick_group(t, lo, hi)
            for gs in self._gstreams:
                cur.wait_stream(gs)
        else:
            for t in range(self.T):
                self._tick(t)
        if self.direct and getattr(self.packed, "precision", None) == 2:     # the bootstrap values: the two critics only
            self.packed.values(self.cur1, self.cur2, out=(b["last_vf"][:, 0], b["last_vf"][:, 1]))
        else:
            _, v1, _, v2 = self._forward_both(self.cur1, self.cur2)
            b["last_vf"][:, 0], b["last_vf"][:, 1] = v1, v2
        nat.check(nat.lib().hh_gae(self.T, self.env.n_arenas, b["rew"].data_ptr(), b["vf"].data_ptr(),
                                   b["last_vf"].data_ptr(), b["done"].data_ptr(), self.gamma, self.lam,
                                   b["adv"].data_ptr(), b["vtarg"].data_ptr(), st), "hh_gae")
        # CustomCallback.on_postprocess_trajectory (train_hetero.py:120-160): the critic's action columns get
        # the real actions, scaled; VF_PREDS / advantages above were computed with zeros
        if self.native_glue:
            nat.check(nat.lib().hh_fragment_writeback(self.T, self.env.n_arenas, D, b["actions"].data_ptr(), b["flat1"].data_ptr(),
                                                      b["flat2"].data_ptr(), st), "hh_fragment_writeback")
            return
        a = b["actions"].to(torch.float32)
        own1, own2 = a[:, :, 0, :] / self.scale, a[:, :, 1, :3] / self.scale[:3]
        b["flat1"][:, :, 0:4], b["flat1"][:, :, 4:7] = own1, own2
        b["flat2"][:, :, 0:3], b["flat2"][:, :, 3:7] = own2, own1

    @torch.no_grad()
    def collect(self):
        """One rollout fragment: dict of [T, N, ...] CUDA tensors (the sampler's own buffers)."""
        if not self._started:
            obs1, obs2 = self.env.reset()
            self._set_obs(obs1, obs2)
            self._started = True
        prev_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(self.allow_tf32)
        try:
            return self._collect()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev_tf32

    def _collect(self):
        i = self._bi
        self.buf = self.bufs[i]
        self._bi = (i + 1) % len(self.bufs)
        if self._copy_done[i] is not None:       # collect_host(): this set's previous fragment is still on its way to the host
            torch.cuda.current_stream(self.dev).wait_event(self._copy_done[i])
        if not self.use_graph:
            self._fragment()
            return self.buf
        if self._graphs[i] is None:
            s = torch.cuda.Stream(self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):
                self._fragment()                 # warm-up (allocations, cuBLAS handles) outside capture
            torch.cuda.current_stream(self.dev).wait_stream(s)
            self._graphs[i] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graphs[i]):
                self._fragment()
            return self.buf                       # the captured region ran once during capture-free warm-up
        self._graphs[i].replay()
        return self.buf

    @torch.no_grad()
    def collect_host(self):
        """collect() + the copy of the whole fragment batch into pinned host memory (what a host-side learner or RLlib's train
        batch reads).  Returns (host dict, event): the copy runs on its own stream and `event` completes when the dict holds this
        fragment.  With n_buffers = 2 the next collect_host() samples into the other buffer set while this copy is in flight
        (host buffers alternate likewise): read the dict after event.synchronize() and before the call after next."""
        b = self.collect()
        i = (self._bi - 1) % len(self.bufs)
        if self._host is None:
            self._host = [{k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in bb.items()} for bb in self.bufs]
            self._copy_stream = torch.cuda.Stream(self.dev)
        cs = self._copy_stream
        cs.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(cs):
            for k, v in b.items():
                self._host[i][k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        self._copy_done[i] = ev
        return self._host[i], ev

    @property
    def env_steps_per_fragment(self):
        return self.T * self.env.n_arenas
