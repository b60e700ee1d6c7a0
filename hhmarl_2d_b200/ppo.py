"""PPO learner for the two heterogeneous policies (ac1_policy, ac2_policy), all on the device.

Hyper-parameters follow train_hetero.py:212-217 (gamma 0.99, lambda 0.95, clip 0.25, lr 1e-4, kl_target 0.025)
and RLlib 2.4's PPO defaults recalled in SURVEY.md Appendix C (kl_coeff 0.2 adaptive x1.5 / x0.5, vf_clip 10,
vf_loss_coeff 1, entropy_coeff 0, max_seq_len 20, advantages standardised per policy batch).  RLlib is absent in
this environment, so these semantics are restated, not pinned ("parity unpinned").

Deviations (deliberate, documented in DESIGN.md): one Adam over the union of both policies' parameters with
loss = loss_ac1 + loss_ac2 (RLlib keeps one optimiser per policy, each containing the process-wide SHARED_LAYER);
rollouts are fixed-length fragments cut every `max_seq_len` ticks rather than per-episode sequences.
Multi-GPU: arenas are sharded, gradients are averaged with ONE all-reduce per minibatch over the flat gradient buffer
of both policies; advantage statistics and the KL statistic are all-reduced.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch.func import functional_call

from .sampler import multicategorical_logp_entropy_kl


class _FlatForward(torch.nn.Module):
    """forward_flat of a policy model as a Module.forward, so that torch.func.functional_call can re-bind its parameters."""

    def __init__(self, model):
        super().__init__()
        self.m = model

    def forward(self, flat, seq_lens):
        return self.m.forward_flat(flat, seq_lens)


class _FusedPPOLoss(torch.autograd.Function):
    """One policy's PPO loss from its per-row terms as ONE kernel (hh_ppo_loss): returns [mean loss, mean kl, mean value loss,
    mean entropy]; the backward is one multiply of the stored per-row derivatives.  CUDA float32 only."""

    @staticmethod
    def forward(ctx, logp, ent, kl, vf, old_logp, adv, vtarg, kl_coeff, clip, vf_clip, vf_coeff, ent_coeff):
        from . import _native as nat
        n = logp.shape[0]
        args = [t.contiguous() for t in (logp, ent, kl, vf, old_logp, adv, vtarg)]
        sums = torch.empty(4, dtype=torch.float32, device=logp.device)
        deriv = torch.empty((4, n), dtype=torch.float32, device=logp.device)
        st = torch.cuda.current_stream(logp.device).cuda_stream
        nat.check(nat.lib().hh_ppo_loss(n, *[t.data_ptr() for t in args], kl_coeff.data_ptr(), float(clip), float(vf_clip),
                                        float(vf_coeff), float(ent_coeff), sums.data_ptr(), deriv.data_ptr(), st), "hh_ppo_loss")
        ctx.save_for_backward(deriv)
        return sums / n

    @staticmethod
    def backward(ctx, g):
        (deriv,) = ctx.saved_tensors
        d = deriv * g[0]                     # only the loss (element 0) is differentiated; the other three are statistics
        return d[0], d[1], d[2], d[3], None, None, None, None, None, None, None, None


def _seq_major(x, L):
    """[T, N, ...] -> [(T/L)*N*L, ...] ordered (chunk, arena, time): rows of one sequence are contiguous."""
    T, N = x.shape[0], x.shape[1]
    C = T // L
    return x.reshape(C, L, N, *x.shape[2:]).transpose(1, 2).reshape(C * N * L, *x.shape[2:])


class PPOLearner:
    """Device-resident learner: parameters and gradients of both policies live in ONE flat buffer each (the modules'
    tensors are views into them), so the gradient exchange is a single all-reduce on the flat gradient with no
    pack / unpack, the optimiser is one fused Adam over one tensor, shuffling is a device randperm and the statistics
    stay on the device until the end of update() (one host synchronisation per update)."""

    def __init__(self, model1, model2, lr=1e-4, clip_param=0.25, kl_target=0.025, kl_coeff=0.2, vf_clip_param=10.0,
                 vf_loss_coeff=1.0, entropy_coeff=0.0, num_sgd_iter=30, sgd_minibatch_size=256, max_seq_len=20,
                 seed=0, use_cuda_graph=True, matmul_tf32=False):
        self.models = (model1, model2)
        self.splits = ((13, 9, 2, 2), (13, 9, 2))
        seen, params = set(), []
        for m in self.models:
            for p in m.parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    params.append(p)
        self.params = params
        dev = params[0].device
        n_tot = sum(p.numel() for p in params)
        self.flat = torch.zeros(n_tot, device=dev, requires_grad=True)       # the leaf the optimiser owns
        self.flat.grad = torch.zeros(n_tot, device=dev)
        o = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                self.flat[o:o + n].copy_(p.reshape(-1))
                p.data = self.flat.data[o:o + n].view_as(p)                  # the module now reads the flat buffer
                p.grad = self.flat.grad[o:o + n].view_as(p)                  # autograd accumulates into the flat gradient
                o += n
        self.grad_bytes = 4 * n_tot
        # The loss is differentiated with respect to `flat` itself: each step re-binds the modules' parameter names to
        # differentiable views flat.split(...) (one cat in the backward = the flat gradient).  `flat` is then the only leaf of
        # the step's autograd graph, so no reference a caller may hold to a graph of the modules' own Parameters (a stray
        # weight.clone(), a cached value output) can pin an AccumulateGrad node to another stream -- which breaks a capture.
        self._sizes = [p.numel() for p in params]
        index = {id(p): k for k, p in enumerate(params)}
        self._names = [[("m." + name, index[id(p)]) for name, p in m.named_parameters()] for m in self.models]
        self._wrapped = [_FlatForward(m) for m in self.models]
        # On the GPU a minibatch step (gather, both policies' forward, loss, backward, Adam) is ~600 small kernels whose launch
        # cost, not their run time, bounds the update: it is captured once per minibatch size into a CUDA graph and replayed
        # (with more than one rank: forward / backward and optimiser as two graphs around the eager all-reduce).
        # matmul_tf32: run the learner's GEMMs (half of its GPU time as SIMT fp32 SGEMMs) as TF32 tensor-core products.  Off by
        # default -- RLlib's torch learner multiplies in fp32 (torch.backends.cuda.matmul.allow_tf32 is False by default).
        self.matmul_tf32 = bool(matmul_tf32)
        self.use_graph = bool(use_cuda_graph) and dev.type == "cuda"
        try:
            self.opt = torch.optim.Adam([self.flat], lr=lr, fused=dev.type == "cuda", capturable=self.use_graph)
        except (TypeError, RuntimeError):
            self.opt = torch.optim.Adam([self.flat], lr=lr)
            self.use_graph = False
        self._graphs, self._static, self._eager_runs, self._side = {}, None, {}, None
        self.graph_warmup = 2                                                # eager minibatches of a size before its capture
        self.clip, self.kl_target, self.kl_coeff = clip_param, kl_target, [kl_coeff, kl_coeff]
        self.kl_coeff_t = torch.tensor(self.kl_coeff, dtype=torch.float32, device=dev)   # what the loss reads (graph-safe)
        self.vf_clip, self.vf_coeff, self.ent_coeff = vf_clip_param, vf_loss_coeff, entropy_coeff
        self.num_sgd_iter, self.mb, self.L = num_sgd_iter, sgd_minibatch_size, max_seq_len
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.epoch = 0
        self._ar_events = []

    # ------------------------------------------------------------------ gradient exchange (one bucket, no copies)
    def _allreduce_grads(self):
        if self.world == 1:
            return
        timed = getattr(self, "time_allreduce", False) and self.flat.is_cuda
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        dist.all_reduce(self.flat.grad)          # the ONE collective of the data path (NCCL over NVLink)
        if timed:
            e1.record()
            self._ar_events.append((e0, e1))
        self.flat.grad.div_(self.world)

    _ar_events: list = []

    def allreduce_us(self):
        """Mean device time of the gradient all-reduce per minibatch since the last call (time_allreduce = True), or None."""
        if not self._ar_events:
            return None
        torch.cuda.synchronize()
        v = [a.elapsed_time(b) * 1e3 for a, b in self._ar_events]
        self._ar_events = []
        return sum(v) / len(v)

    def _global_mean_std(self, x):
        s = torch.stack([x.sum(), (x * x).sum(), torch.tensor(float(x.numel()), device=x.device)])
        if self.world > 1:
            dist.all_reduce(s)
        mean = s[0] / s[2]
        var = (s[1] / s[2] - mean * mean).clamp_min(0.0)
        return mean, var.sqrt()

    def _loss(self, i, views, flat, actions, old_logits, old_logp, adv, vtarg, seq_lens):
        logits, vf = functional_call(self._wrapped[i], {name: views[k] for name, k in self._names[i]}, (flat, seq_lens))
        logp, ent, kl = multicategorical_logp_entropy_kl(logits, actions, self.splits[i], old_logits)
        if logp.is_cuda and logp.dtype == torch.float32:
            out = _FusedPPOLoss.apply(logp, ent, kl, vf, old_logp, adv, vtarg, self.kl_coeff_t[i:i + 1], self.clip, self.vf_clip,
                                      self.vf_coeff, self.ent_coeff)
            return out[0], out[1].detach(), out[2].detach(), out[3].detach()
        ratio = torch.exp(logp - old_logp)
        surr = torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - self.clip, 1 + self.clip))
        vf_loss = torch.clamp((vf - vtarg) ** 2, 0, self.vf_clip)
        loss = (-surr + self.kl_coeff_t[i] * kl + self.vf_coeff * vf_loss - self.ent_coeff * ent).mean()
        return loss, kl.mean().detach(), vf_loss.mean().detach(), ent.mean().detach()

    def update(self, batch: dict, num_sgd_iter: int | None = None):
        """batch: the VecSampler buffers ([T, N, ...]).  Returns a dict of statistics."""
        L = self.L
        T = batch["rew"].shape[0]
        assert T % L == 0, "fragment length must be a multiple of max_seq_len"
        data = []
        for i in range(2):
            adv = _seq_major(batch["adv"][:, :, i], L)
            mean, std = self._global_mean_std(adv)
            data.append(dict(flat=_seq_major(batch["flat1" if i == 0 else "flat2"], L),
                             actions=_seq_major(batch["actions"][:, :, i, :len(self.splits[i])], L),
                             logits=_seq_major(batch["logits1" if i == 0 else "logits2"], L),
                             logp=_seq_major(batch["logp"][:, :, i], L),
                             adv=(adv - mean) / std.clamp_min(1e-4), vtarg=_seq_major(batch["vtarg"][:, :, i], L)))
        dev = data[0]["flat"].device
        if self.use_graph:
            data = self._to_static(data)
        n_seq = data[0]["flat"].shape[0] // L
        seq_per_mb = max(1, min(n_seq, self.mb // L))     # a batch smaller than one minibatch is ONE (smaller) minibatch
        iters = self.num_sgd_iter if num_sgd_iter is None else num_sgd_iter
        acc = torch.zeros(7, device=dev)                  # loss, kl x2, vf_loss x2, entropy x2: summed on the device
        n_mb = 0
        ar = torch.arange(L, device=dev)
        for _ in range(iters):
            perm = torch.randperm(n_seq, generator=self.gen, device=dev)
            for s in range(0, n_seq, seq_per_mb):         # the last minibatch holds the remainder
                sel = perm[s:s + seq_per_mb]
                rows = (sel[:, None] * L + ar[None, :]).reshape(-1)
                self._minibatch(data, rows, int(sel.shape[0]), acc)
                n_mb += 1
        if self.world > 1 and n_mb:
            kl_sum = acc[1:3].clone()
            dist.all_reduce(kl_sum)
            acc[1:3] = kl_sum / self.world
        vals = (acc / max(1, n_mb)).tolist()              # the one host synchronisation of the update
        stats = {"loss": vals[0], "kl": vals[1:3], "vf_loss": vals[3:5], "entropy": vals[5:7], "minibatches": n_mb}
        if n_mb:                                          # RLlib's adaptive KL coefficient (update_kl)
            for i in range(2):
                if stats["kl"][i] > 2.0 * self.kl_target:
                    self.kl_coeff[i] *= 1.5
                elif stats["kl"][i] < 0.5 * self.kl_target:
                    self.kl_coeff[i] *= 0.5
            self.kl_coeff_t.copy_(torch.tensor(self.kl_coeff, dtype=torch.float32))
        stats["kl_coeff"] = list(self.kl_coeff)
        self.epoch += 1
        return stats

    # ------------------------------------------------------------------ one minibatch: eager, or a CUDA-graph replay
    def _fwd_bwd(self, data, rows, n_sel):
        """Both policies' losses on the rows `rows`, gradient into the flat buffer.  Returns the 7 statistics as one tensor."""
        if self.flat.is_cuda:
            prev = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = self.matmul_tf32
            try:
                return self._fwd_bwd_impl(data, rows, n_sel)
            finally:
                torch.backends.cuda.matmul.allow_tf32 = prev
        return self._fwd_bwd_impl(data, rows, n_sel)

    def _fwd_bwd_impl(self, data, rows, n_sel):
        seq_lens = [self.L] * n_sel
        total, parts = 0.0, []
        views = [v.view(p.shape) for v, p in zip(self.flat.split(self._sizes), self.params)]
        for i in range(2):
            d = data[i]
            loss, kl, vfl, ent = self._loss(i, views, d["flat"][rows], d["actions"][rows], d["logits"][rows], d["logp"][rows],
                                            d["adv"][rows], d["vtarg"][rows], seq_lens)
            total = total + loss
            parts += [kl, vfl, ent]
        self.flat.grad.zero_()
        total.backward()
        return torch.stack([total.detach(), parts[0], parts[3], parts[1], parts[4], parts[2], parts[5]])

    def _minibatch(self, data, rows, n_sel, acc):
        if not self.use_graph:
            st = self._fwd_bwd(data, rows, n_sel)
            self._allreduce_grads()
            self.opt.step()
            acc += st
            return
        g = self._graphs.get(n_sel)
        if g is None:
            runs = self._eager_runs.get(n_sel, 0)
            if runs < self.graph_warmup:       # real steps, on the stream the capture will use (what torch asks for before one)
                if self._side is None:
                    self._side = torch.cuda.Stream()
                side = self._side
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    st = self._fwd_bwd(data, rows, n_sel)
                    self._allreduce_grads()
                    self.opt.step()
                    acc += st
                torch.cuda.current_stream().wait_stream(side)
                self._eager_runs[n_sel] = runs + 1
                return
            g = self._capture(data, n_sel)
            self._graphs[n_sel] = g
        g["rows"].copy_(rows)
        if self.world == 1:
            g["g"].replay()
        else:
            g["g"].replay()
            self._allreduce_grads()
            g["g2"].replay()
        acc += g["stats"]

    def _capture(self, data, n_sel):
        rows = torch.zeros(n_sel * self.L, dtype=torch.int64, device=self.flat.device)
        g = {"rows": rows, "g": torch.cuda.CUDAGraph()}
        if self._side is None:
            self._side = torch.cuda.Stream()
        for m in self.models:                  # drop the value head's cached output (it keeps the last eager step's graph alive)
            if hasattr(m, "_val"):
                m._val = None
        torch.cuda.synchronize()
        if self.world == 1:
            with torch.cuda.graph(g["g"], stream=self._side):
                g["stats"] = self._fwd_bwd(data, rows, n_sel)
                self.opt.step()
        else:
            with torch.cuda.graph(g["g"], stream=self._side):
                g["stats"] = self._fwd_bwd(data, rows, n_sel)
            g["g2"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g["g2"], pool=g["g"].pool(), stream=self._side):
                self.opt.step()
        return g

    def _to_static(self, data):
        """The graphs read the batch at fixed addresses: copy this update's batch into buffers that stay."""
        shapes = [{k: (tuple(v.shape), v.dtype) for k, v in d.items()} for d in data]
        if self._static is None or self._static[0] != shapes:
            self._static = (shapes, [{k: torch.empty_like(v) for k, v in d.items()} for d in data])
            self._graphs, self._eager_runs = {}, {}
        for d, sd in zip(data, self._static[1]):
            for k, v in d.items():
                sd[k].copy_(v)
        return self._static[1]

    # ------------------------------------------------------------------ training state (resume): what algo.save() keeps
    # beyond the policy weights -- optimiser moments, adaptive KL coefficients, epoch counter, shuffling RNG
    def state_dict(self):
        return {"flat": self.flat.detach().clone(), "opt": self.opt.state_dict(), "kl_coeff": list(self.kl_coeff),
                "epoch": self.epoch, "gen": self.gen.get_state()}

    def load_state_dict(self, sd):
        with torch.no_grad():
            self.flat.copy_(sd["flat"].to(self.flat.device))
        self.opt.load_state_dict(sd["opt"])
        self.kl_coeff = list(sd["kl_coeff"])
        self.kl_coeff_t.copy_(torch.tensor(self.kl_coeff, dtype=torch.float32))
        self._graphs, self._eager_runs = {}, {}        # the optimiser's state tensors were replaced: capture again
        self.epoch = int(sd["epoch"])
        self.gen.set_state(sd["gen"].cpu() if self.gen.device.type == "cpu" else sd["gen"])
