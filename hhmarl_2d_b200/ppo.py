"""PPO learner for the two heterogeneous policies (ac1_policy, ac2_policy), all on the device.

Hyper-parameters follow train_hetero.py:212-217 (gamma 0.99, lambda 0.95, clip 0.25, lr 1e-4, kl_target 0.025)
and RLlib 2.4's PPO defaults recalled in SURVEY.md Appendix C (kl_coeff 0.2 adaptive x1.5 / x0.5, vf_clip 10,
vf_loss_coeff 1, entropy_coeff 0, max_seq_len 20, advantages standardised per policy batch).  RLlib is absent in
this environment, so these semantics are restated, not pinned ("parity unpinned").

Deviations (deliberate, documented in DESIGN.md): one Adam over the union of both policies' parameters with
loss = loss_ac1 + loss_ac2 (RLlib keeps one optimiser per policy, each containing the process-wide SHARED_LAYER);
rollouts are fixed-length fragments cut every `max_seq_len` ticks rather than per-episode sequences.
Multi-GPU: arenas are sharded, gradients are averaged with ONE all-reduce per minibatch over a flat bucket that
holds both policies' gradients; advantage statistics and the KL statistic are all-reduced.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .sampler import multicategorical_logp_entropy_kl


def _seq_major(x, L):
    """[T, N, ...] -> [(T/L)*N*L, ...] ordered (chunk, arena, time): rows of one sequence are contiguous."""
    T, N = x.shape[0], x.shape[1]
    C = T // L
    return x.reshape(C, L, N, *x.shape[2:]).transpose(1, 2).reshape(C * N * L, *x.shape[2:])


class PPOLearner:
    def __init__(self, model1, model2, lr=1e-4, clip_param=0.25, kl_target=0.025, kl_coeff=0.2, vf_clip_param=10.0,
                 vf_loss_coeff=1.0, entropy_coeff=0.0, num_sgd_iter=30, sgd_minibatch_size=256, max_seq_len=20,
                 seed=0):
        self.models = (model1, model2)
        self.splits = ((13, 9, 2, 2), (13, 9, 2))
        seen, params = set(), []
        for m in self.models:
            for p in m.parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    params.append(p)
        self.params = params
        self.opt = torch.optim.Adam(params, lr=lr)
        self.clip, self.kl_target, self.kl_coeff = clip_param, kl_target, [kl_coeff, kl_coeff]
        self.vf_clip, self.vf_coeff, self.ent_coeff = vf_clip_param, vf_loss_coeff, entropy_coeff
        self.num_sgd_iter, self.mb, self.L = num_sgd_iter, sgd_minibatch_size, max_seq_len
        self.gen = torch.Generator(device="cpu")
        self.gen.manual_seed(seed)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._flat = None

    # ------------------------------------------------------------------ gradient exchange (one bucket)
    def _allreduce_grads(self):
        if self.world == 1:
            return
        if self._flat is None:
            self._flat = torch.zeros(sum(p.numel() for p in self.params), device=self.params[0].device)
        o = 0
        for p in self.params:
            n = p.numel()
            self._flat[o:o + n] = p.grad.reshape(-1) if p.grad is not None else 0
            o += n
        dist.all_reduce(self._flat)
        self._flat /= self.world
        o = 0
        for p in self.params:
            n = p.numel()
            if p.grad is not None:
                p.grad.copy_(self._flat[o:o + n].view_as(p.grad))
            o += n

    def _global_mean_std(self, x):
        s = torch.stack([x.sum(), (x * x).sum(), torch.tensor(float(x.numel()), device=x.device)])
        if self.world > 1:
            dist.all_reduce(s)
        mean = s[0] / s[2]
        var = (s[1] / s[2] - mean * mean).clamp_min(0.0)
        return mean, var.sqrt()

    def _loss(self, i, flat, actions, old_logits, old_logp, adv, vtarg, seq_lens):
        logits, vf = self.models[i].forward_flat(flat, seq_lens)
        logp, ent, kl = multicategorical_logp_entropy_kl(logits, actions, self.splits[i], old_logits)
        ratio = torch.exp(logp - old_logp)
        surr = torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - self.clip, 1 + self.clip))
        vf_loss = torch.clamp((vf - vtarg) ** 2, 0, self.vf_clip)
        loss = (-surr + self.kl_coeff[i] * kl + self.vf_coeff * vf_loss - self.ent_coeff * ent).mean()
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
        n_seq = data[0]["flat"].shape[0] // L
        seq_per_mb = max(1, self.mb // L)
        iters = self.num_sgd_iter if num_sgd_iter is None else num_sgd_iter
        stats = {"loss": 0.0, "kl": [0.0, 0.0], "vf_loss": [0.0, 0.0], "entropy": [0.0, 0.0], "minibatches": 0}
        ar = torch.arange(L, device=data[0]["flat"].device)
        for _ in range(iters):
            perm = torch.randperm(n_seq, generator=self.gen).to(ar.device)
            for s in range(0, n_seq - seq_per_mb + 1, seq_per_mb):
                rows = (perm[s:s + seq_per_mb, None] * L + ar[None, :]).reshape(-1)
                seq_lens = torch.full((seq_per_mb,), L, dtype=torch.int32)
                total = 0.0
                for i in range(2):
                    d = data[i]
                    loss, kl, vfl, ent = self._loss(i, d["flat"][rows], d["actions"][rows], d["logits"][rows],
                                                    d["logp"][rows], d["adv"][rows], d["vtarg"][rows], seq_lens)
                    total = total + loss
                    stats["kl"][i] += float(kl)
                    stats["vf_loss"][i] += float(vfl)
                    stats["entropy"][i] += float(ent)
                self.opt.zero_grad(set_to_none=False)
                total.backward()
                self._allreduce_grads()
                self.opt.step()
                stats["loss"] += float(total.detach())
                stats["minibatches"] += 1
        m = max(1, stats["minibatches"])
        for k in ("kl", "vf_loss", "entropy"):
            stats[k] = [v / m for v in stats[k]]
        stats["loss"] /= m
        for i in range(2):   # RLlib's adaptive KL coefficient (update_kl)
            kl = torch.tensor(stats["kl"][i], device=ar.device)
            if self.world > 1:
                dist.all_reduce(kl)
                kl /= self.world
            if float(kl) > 2.0 * self.kl_target:
                self.kl_coeff[i] *= 1.5
            elif float(kl) < 0.5 * self.kl_target:
                self.kl_coeff[i] *= 0.5
        stats["kl_coeff"] = list(self.kl_coeff)
        return stats
