"""Packed forward of BOTH sampler policies (Fight1 + Fight2 or Esc1 + Esc2, actor + central critic) for the
one-token-per-sequence case of the rollout (seq_lens = [1] * B), as a handful of large GEMMs instead of ~60
small ones.  Pure re-association of the same fp32 arithmetic as models.py:

  stage 1  H_p = tanh(x_p @ W1_p + b1_p)          every input branch of policy p as ONE [57|66 -> 1000|1000] GEMM
           (block-sparse weight: a branch only sees its own slice of the flattened central observation)
  stage 2  attention over a single token = out_proj(v_proj(.)): folded into one matrix per block, residual +
           L2-normalise (Fight only)
  stage 3  the process-wide SHARED_LAYER (ac_models_hetero.py:22-27) is the same 500x500 matrix for all four
           chains (2 policies x actor/critic): ONE [4B, 500] x [500, 500] GEMM
  stage 4  heads: logits (26 / 24) and values

`refresh()` re-packs after the learner changed the weights.  Parity with models.forward_flat is tested in
tests/test_models.py (CPU) to ~1e-6.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import models as M


def _lin(fc):
    l = fc._model[0]
    return l.weight, l.bias


class PackedPolicyPair:
    _LISTS = ("W1", "b1", "Watt", "batt", "Wact", "bact", "Wval", "bval")

    def __init__(self, model1, model2):
        self.m = (model1, model2)
        self.fight = isinstance(model1, M.Fight1)
        self._pack()

    @torch.no_grad()
    def refresh(self):
        """Re-pack IN PLACE (the packed tensors may be baked into a captured CUDA graph)."""
        old = {k: getattr(self, k) for k in self._LISTS + ("Ws", "bs")}
        self._pack()
        for k in self._LISTS:
            for dst, src in zip(old[k], getattr(self, k)):
                dst.copy_(src)
            setattr(self, k, old[k])
        old["Ws"].copy_(self.Ws)
        old["bs"].copy_(self.bs)
        self.Ws, self.bs = old["Ws"], old["bs"]

    @torch.no_grad()
    def _pack(self):
        dev = next(self.m[0].parameters()).device
        self.W1, self.b1, self.Watt, self.batt, self.Wact, self.bact, self.Wval, self.bval = [], [], [], [], [], [], [], []
        for m in self.m:
            a, b = m.own_act_dim, m.other_act_dim
            own, oth = m.own_obs_dim, m.other_obs_dim
            D = a + b + own + oth
            o_own, o_oth = a + b, a + b + own          # flat = [act_own | act_oth | obs_own | obs_oth]
            W = torch.zeros((D, 1000), device=dev)
            bias = torch.zeros(1000, device=dev)

            def put(fc, cols, rows):
                w, bb = _lin(fc)                        # w: [out, in] over the concatenation of `rows`
                c0, c1 = cols
                k = 0
                for r0, r1 in rows:
                    W[r0:r1, c0:c1] = w[:, k:k + (r1 - r0)].t()
                    k += r1 - r0
                bias[c0:c1] = bb

            if self.fight:
                ss = m.ss
                put(m.inp1, (0, 200), [(o_own, o_own + ss)])
                put(m.inp2, (200, 400), [(o_own + ss, o_own + own)])
                put(m.inp3, (400, 500), [(o_own, o_own + own)])
                put(m.v1, (500, 675), [(o_own, o_own + own), (0, a)])
                put(m.v2, (675, 850), [(o_oth, o_oth + oth), (a, a + b)])
                put(m.v3, (850, 1000), [(o_own, o_own + own), (0, a), (o_oth, o_oth + oth), (a, a + b)])
                # single-token attention: out_proj(v_proj(x)) = x @ (Wo Wv)^T + (Wo bv + bo)
                Wa = torch.zeros((250, 250), device=dev)
                ba = torch.zeros(250, device=dev)
                for mha, (c0, c1) in ((m.att_act, (0, 100)), (m.att_val, (100, 250))):
                    e = mha.embed_dim
                    Wv, bv = mha.in_proj_weight[2 * e:], mha.in_proj_bias[2 * e:]
                    Wo, bo = mha.out_proj.weight, mha.out_proj.bias
                    Wa[c0:c1, c0:c1] = (Wo @ Wv).t()
                    ba[c0:c1] = Wo @ bv + bo
                self.Watt.append(Wa)
                self.batt.append(ba)
            else:
                k1 = m.k1
                put(m.inp1, (0, 150), [(o_own, o_own + k1)])
                put(m.inp2, (150, 400), [(o_own + k1, o_own + k1 + 18)])
                put(m.inp3, (400, 500), [(o_own + k1 + 18, o_own + own)])
                put(m.inp1_val, (500, 1000), [(o_own, o_own + own), (0, a), (o_oth, o_oth + oth), (a, a + b)])
            self.W1.append(W)
            self.b1.append(bias)
            wa, ba_ = _lin(m.act_out)
            wv, bv_ = _lin(m.val_out)
            self.Wact.append(wa.t().contiguous())
            self.bact.append(ba_.clone())
            self.Wval.append(wv.t().contiguous())
            self.bval.append(bv_.clone())
        ws, bs = _lin(self.m[0].shared_layer)
        self.Ws, self.bs = ws.t().contiguous(), bs.clone()

    @torch.no_grad()
    def forward(self, flat1, flat2):
        """-> (logits1 [B,26], value1 [B], logits2 [B,24], value2 [B])"""
        B = flat1.shape[0]
        stacked = torch.empty((4 * B, 500), device=flat1.device, dtype=flat1.dtype)
        for p, x in enumerate((flat1, flat2)):
            H = torch.tanh(torch.addmm(self.b1[p], x, self.W1[p]))
            act_in, val_in = stacked[2 * p * B:(2 * p + 1) * B], stacked[(2 * p + 1) * B:(2 * p + 2) * B]
            if self.fight:
                full = torch.cat((H[:, 400:500], H[:, 850:1000]), dim=1)
                att = torch.addmm(self.batt[p], full, self.Watt[p])
                r = full + att
                act_in[:, :400] = H[:, :400]
                act_in[:, 400:] = F.normalize(r[:, :100])
                val_in[:, :350] = H[:, 500:850]
                val_in[:, 350:] = F.normalize(r[:, 100:])
            else:
                act_in.copy_(H[:, :500])
                val_in.copy_(H[:, 500:])
        Z = torch.tanh(torch.addmm(self.bs, stacked, self.Ws))
        l1 = torch.addmm(self.bact[0], Z[0:B], self.Wact[0])
        v1 = torch.addmm(self.bval[0], Z[B:2 * B], self.Wval[0])[:, 0]
        l2 = torch.addmm(self.bact[1], Z[2 * B:3 * B], self.Wact[1])
        v2 = torch.addmm(self.bval[1], Z[3 * B:], self.Wval[1])[:, 0]
        return l1, v1, l2, v2
