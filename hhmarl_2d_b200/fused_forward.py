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

import ctypes

import torch
import torch.nn.functional as F

from . import _native as nat
from . import models as M


def _lin(fc):
    l = fc._model[0]
    return l.weight, l.bias


class PackedPolicyPair:
    _LISTS = ("W1", "b1", "Watt", "batt", "Wact", "bact", "Wval", "bval")

    def __init__(self, model1, model2):
        # one SHARED_LAYER per process (ac_models_hetero.py:22-27): the packed forward applies model1's to all four chains
        if model1.shared_layer is not model2.shared_layer:
            raise ValueError("PackedPolicyPair needs two policies that share ONE shared_layer module "
                             "(models.build_policy_pair); these were built separately")
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


def _pack_image(w_rm: torch.Tensor, n_cols: int, n_total: int, n_chunk: int, row_shift: int, ksteps: int, kps: int = 1,
                img=None, us=None):
    """Operand image of a zero-padded row-major [K, ld] fp32 weight matrix for the tcgen05 path (hh_policy_pack):
    fp16 hi / lo halves of 2^s w in the K-major canonical shared-memory layout, one ring stage after the other in
    the order the kernel consumes them (for the CTA-pair kernel each stage is split into the two CTAs' column halves).
    -> (image uint8 tensor, scale pair float[2]); re-packs IN PLACE when given."""
    L = nat.lib()
    dev = w_rm.device
    if img is None:
        img = torch.empty(int(L.hh_policy_image_bytes(ksteps, n_total)), dtype=torch.uint8, device=dev)
        us = torch.zeros(2, dtype=torch.float32, device=dev)
    assert w_rm.is_contiguous() and w_rm.dtype == torch.float32 and w_rm.is_cuda
    st = torch.cuda.current_stream(dev).cuda_stream
    nat.check(L.hh_policy_pack(w_rm.data_ptr(), w_rm.shape[0], min(n_cols, w_rm.shape[1]), w_rm.stride(0), n_total, n_chunk, row_shift,
                               ksteps, kps, img.data_ptr(), us.data_ptr(), st), "hh_policy_pack")
    return img, us


def _att_image_geometry(att_lo: int, att_n: int):
    """The attention block reads activation columns [att_lo, att_lo + att_n); the MMA's A operand has to start on a core
    matrix (8 columns), so the image's K range starts at floor8(att_lo) and its first rows are zero; its N is the block
    padded to the MMA's N granularity (16 columns for the CTA-pair kernel, 8 otherwise)."""
    mode = nat.lib().hh_policy_tc_mode()              # 0: 64-row tiles, 1: CTA pairs, 2: 128-row tiles (A_lo in tensor memory)
    k0 = att_lo & ~15 if mode == 2 else att_lo & ~7   # the TMEM operand of the 128-row form starts on a K = 16 step
    g = 16 if mode >= 1 else 8
    return att_lo - k0, (att_lo + att_n - k0 + 15) // 16, (att_n + g - 1) // g * g          # (row_shift, ksteps, N)


class FusedPolicyPair:
    """The same forward as PackedPolicyPair.forward in ONE launch of the hand-written kernel csrc/hh_policy.cu
    (C ABI hh_policy_forward): per 64-row tile and chain the activations stay in shared memory, every GEMM runs on
    the tensor cores as 3xTF32 (fp32-equivalent, precision=0) or plain TF32 (precision=1), both on the legacy mma.sync
    path, or -- precision=2, csrc/hh_policy_tc.cu -- on tcgen05 with TMEM accumulators (fp16 hi / lo operand split,
    fp32-equivalent).  Weights are re-packed into the layouts the kernels expect; `refresh()` re-packs in place
    (CUDA-graph safe)."""

    NP, NW, NH = 504, 512, 32     # K extent / N extent (row stride) of the 500-wide layers, head width

    def __init__(self, model1, model2, precision: int = 0):
        self.packed = PackedPolicyPair(model1, model2)
        self.fight = self.packed.fight
        self.precision = int(precision)
        self.dev = next(model1.parameters()).device
        self._alloc()
        self._fill()
        self._out = None

    def _alloc(self):
        z = lambda *shape: torch.zeros(shape, device=self.dev, dtype=torch.float32)  # noqa: E731
        self.k1_pad = [((w.shape[0] + 7) // 8) * 8 for w in self.packed.W1]
        self.w1 = [[z(self.k1_pad[p], self.NW) for _ in range(2)] for p in range(2)]      # [policy][actor, critic]
        self.b1 = [[z(self.NW) for _ in range(2)] for p in range(2)]
        self.att = [(400, 100, 104), (350, 150, 152)]                                     # (att_lo, att_n, att_pad)
        self.watt = [[z(self.att[k][2], self.att[k][2]) for k in range(2)] for p in range(2)] if self.fight else None
        self.batt = [[z(self.att[k][2]) for k in range(2)] for p in range(2)] if self.fight else None
        self.wh = [[z(self.NP, self.NH) for _ in range(2)] for p in range(2)]
        self.bh = [[z(self.NH) for _ in range(2)] for p in range(2)]
        self.ws, self.bs = z(self.NP, self.NW), z(self.NW)

    @staticmethod
    def _to_fragments(dst: torch.Tensor, w: torch.Tensor):
        """Row-major zero-padded [K, N] -> MMA fragment order [K/8][N/8][lane = 4 g + t][(b0, b1)] with
        b0 = w[8 ks + t, 8 nt + g], b1 = w[8 ks + t + 4, 8 nt + g] (in place into dst, same number of elements)."""
        K, N = w.shape
        v = w.reshape(K // 8, 2, 4, N // 8, 8).permute(0, 3, 4, 2, 1)      # (ks, q, t, nt, g) -> (ks, nt, g, t, q)
        dst.view(-1).copy_(v.reshape(-1))

    @torch.no_grad()
    def _fill(self):
        self._fill_row_major()
        for p in range(2):
            for k in range(2):
                self._to_fragments(self.w1[p][k], self._rm["w1"][p][k])
                self._to_fragments(self.wh[p][k], self._rm["wh"][p][k])
                if self.fight:
                    self._to_fragments(self.watt[p][k], self._rm["watt"][p][k])
        self._to_fragments(self.ws, self._rm["ws"])
        if self.precision == 2:
            self._pack_images()

    def _pack_images(self):
        """tcgen05 path: operand images of every weight matrix (in place after the first call)."""
        first = not hasattr(self, "img")
        if first:
            self.img = {}
        rm = self._rm

        def put(key, w, n_cols, n_total, n_chunk, shift, ksteps, kps=1):
            old = self.img.get(key, (None, None))
            self.img[key] = _pack_image(w, n_cols, n_total, n_chunk, shift, ksteps, kps, *old)

        for p in range(2):
            k1s = (self.packed.W1[p].shape[0] + 15) // 16
            for k in range(2):
                put(("w1", p, k), rm["w1"][p][k], 512, 512, 256, 0, k1s)
                put(("wh", p, k), rm["wh"][p][k], 32, 32, 32, 0, 32, 8)
                if self.fight:
                    lo, n, pad = self.att[k]
                    shift, ks, nn = _att_image_geometry(lo, n)
                    put(("watt", p, k), rm["watt"][p][k], n, nn, nn, shift, ks)
        put(("ws",), rm["ws"], 512, 512, 256, 0, 32)

    @torch.no_grad()
    def _fill_row_major(self):
        """Zero-padded row-major copies (staging for _to_fragments; the kernel reads the fragment-ordered tensors)."""
        if not hasattr(self, "_rm"):
            self._rm = {"w1": [[torch.zeros_like(t) for t in row] for row in self.w1],
                        "wh": [[torch.zeros_like(t) for t in row] for row in self.wh],
                        "watt": [[torch.zeros_like(t) for t in row] for row in self.watt] if self.fight else None,
                        "ws": torch.zeros_like(self.ws)}
        rm = self._rm
        pk = self.packed
        for p in range(2):
            D = pk.W1[p].shape[0]
            for k, cols in enumerate((slice(0, 500), slice(500, 1000))):
                rm["w1"][p][k][:D, :500].copy_(pk.W1[p][:, cols])
                self.b1[p][k][:500].copy_(pk.b1[p][cols])
            if self.fight:
                for k, blk in enumerate((slice(0, 100), slice(100, 250))):
                    n = self.att[k][1]
                    rm["watt"][p][k][:n, :n].copy_(pk.Watt[p][blk, blk])
                    self.batt[p][k][:n].copy_(pk.batt[p][blk])
            n_act = pk.Wact[p].shape[1]
            rm["wh"][p][0][:500, :n_act].copy_(pk.Wact[p])
            self.bh[p][0][:n_act].copy_(pk.bact[p])
            rm["wh"][p][1][:500, :1].copy_(pk.Wval[p])
            self.bh[p][1][:1].copy_(pk.bval[p])
        rm["ws"][:500, :500].copy_(pk.Ws)
        self.bs[:500].copy_(pk.bs)

    @torch.no_grad()
    def refresh(self):
        self.packed.refresh()
        self._fill()

    @torch.no_grad()
    def forward(self, flat1, flat2, out=None):
        """-> (logits1 [B,26|..], value1 [B], logits2, value2); `out` = preallocated tuple of the same tensors."""
        B = flat1.shape[0]
        n_act = [self.packed.Wact[p].shape[1] for p in range(2)]
        if out is None:
            if self._out is None or self._out[0].shape[0] != B:
                self._out = (torch.empty((B, n_act[0]), device=self.dev), torch.empty((B,), device=self.dev),
                             torch.empty((B, n_act[1]), device=self.dev), torch.empty((B,), device=self.dev))
            out = self._out
        if self.precision == 2:
            return self._forward_tc(flat1, flat2, out, n_act)
        chains = (nat.HHPolicyChain * 4)()
        for p, x in enumerate((flat1, flat2)):
            assert x.is_contiguous() and x.dtype == torch.float32 and x.is_cuda
            for k in range(2):
                c = chains[2 * p + k]
                c.x, c.w1, c.b1 = x.data_ptr(), self.w1[p][k].data_ptr(), self.b1[p][k].data_ptr()
                if self.fight:
                    c.watt, c.batt = self.watt[p][k].data_ptr(), self.batt[p][k].data_ptr()
                    c.att_lo, c.att_n, c.att_pad = self.att[k]
                else:
                    c.watt, c.batt, c.att_lo, c.att_n, c.att_pad = None, None, 0, 0, 0
                c.wh, c.bh = self.wh[p][k].data_ptr(), self.bh[p][k].data_ptr()
                o = out[2 * p + k]
                c.out, c.ld_out, c.n_out = o.data_ptr(), o.stride(0), (n_act[p] if k == 0 else 1)
                c.ldx, c.d_in, c.k1_pad = x.stride(0), x.shape[1], self.k1_pad[p]
        st = torch.cuda.current_stream(self.dev).cuda_stream
        nat.check(nat.lib().hh_policy_forward(B, chains, self.ws.data_ptr(), self.bs.data_ptr(), self.precision, st),
                  "hh_policy_forward")
        return out


    @torch.no_grad()
    def forward_one(self, p: int, flat, out=None):
        """Actor + central critic of ONE policy (p = 0: ac1_policy, 1: ac2_policy) on its flattened central observation:
        what RLlib's Policy.compute_actions evaluates for one policy's batch.  -> (logits [B, 26|24], value [B])"""
        assert self.precision == 2, "forward_one runs on the tcgen05 path (precision=2)"
        B = flat.shape[0]
        n_act = self.packed.Wact[p].shape[1]
        if out is None:
            out = (torch.empty((B, n_act), device=self.dev), torch.empty((B,), device=self.dev))
        self._launch_tc([(p, flat, out)], [n_act if q == p else None for q in range(2)])
        return out

    def _forward_tc(self, flat1, flat2, out, n_act):
        self._launch_tc([(0, flat1, out[0:2]), (1, flat2, out[2:4])], n_act)
        return out

    def values(self, flat1, flat2, out):
        """Only the two central critics (the bootstrap value at the end of a fragment): half the chains of forward().
        `out` = (value1 [B], value2 [B]).  tcgen05 path only."""
        assert self.precision == 2
        self._launch_tc([(0, flat1, (None, out[0])), (1, flat2, (None, out[1]))], [None, None], kinds=(1,))
        return out

    def _launch_tc(self, jobs, n_act, kinds=(0, 1)):
        nk = len(kinds)
        chains = (nat.HHPolicyChainEx * (nk * len(jobs)))()
        for j, (p, x, o2) in enumerate(jobs):
            B = x.shape[0]
            assert x.dtype == torch.float32 and x.is_cuda and x.stride(1) == 1
            for i, k in enumerate(kinds):
                c = chains[nk * j + i]
                c.x, c.ldx, c.d_in, c.k1_pad = x.data_ptr(), x.stride(0), x.shape[1], self.k1_pad[p]
                c.b1, c.bs, c.bh = self.b1[p][k].data_ptr(), self.bs.data_ptr(), self.bh[p][k].data_ptr()
                c.img_w1, c.us_w1 = (t.data_ptr() for t in self.img[("w1", p, k)])
                c.img_ws, c.us_ws = (t.data_ptr() for t in self.img[("ws",)])
                c.img_wh, c.us_wh = (t.data_ptr() for t in self.img[("wh", p, k)])
                if self.fight:
                    c.batt = self.batt[p][k].data_ptr()
                    c.att_lo, c.att_n, c.att_pad = self.att[k]
                    c.img_att, c.us_att = (t.data_ptr() for t in self.img[("watt", p, k)])
                o = o2[k]
                c.out, c.ld_out, c.n_out = o.data_ptr(), o.stride(0), (n_act[p] if k == 0 else 1)
                c.n_rows = B
        st = torch.cuda.current_stream(self.dev).cuda_stream
        nat.check(nat.lib().hh_policy_forward_ex(len(chains), chains, 2, st), "hh_policy_forward_ex")


class FusedActor:
    """One frozen actor network -- `model.actor(own_obs)` of Fight1/Fight2/Esc1/Esc2 -- packed for the general
    entry point hh_policy_forward_ex: what HHMARLBaseEnv._policy_actions evaluates for a self-play opponent
    (env_base.py:349-398) and for every aircraft inside HighLevelEnv (env_hier.py:114-140)."""

    NP, NW, NH = 504, 512, 32

    def __init__(self, model):
        self.model = model
        self.fight = isinstance(model, M.Fight1)
        self.dev = next(model.parameters()).device
        self.d_in = model.own_obs_dim
        self.k1_pad = ((self.d_in + 7) // 8) * 8
        self.splits = M.ACTION_SPLITS[model.ac_type]
        self.n_out = sum(self.splits)
        z = lambda *shape: torch.zeros(shape, device=self.dev, dtype=torch.float32)  # noqa: E731
        self.w1, self.b1 = z(self.k1_pad, self.NW), z(self.NW)
        self.watt, self.batt = (z(104, 104), z(104)) if self.fight else (None, None)
        self.ws, self.bs = z(self.NP, self.NW), z(self.NW)
        self.wh, self.bh = z(self.NP, self.NH), z(self.NH)
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        m, own = self.model, self.d_in
        W = torch.zeros((self.k1_pad, self.NW), device=self.dev)
        if self.fight:
            ss = m.ss
            blocks = ((m.inp1, 0, 200, 0, ss), (m.inp2, 200, 400, ss, own), (m.inp3, 400, 500, 0, own))
        else:
            k1 = m.k1
            blocks = ((m.inp1, 0, 150, 0, k1), (m.inp2, 150, 400, k1, k1 + 18), (m.inp3, 400, 500, k1 + 18, own))
        self.b1.zero_()
        for fc, c0, c1, r0, r1 in blocks:
            w, b = _lin(fc)
            W[r0:r1, c0:c1] = w.t()
            self.b1[c0:c1] = b
        FusedPolicyPair._to_fragments(self.w1, W)
        img = getattr(self, "img", {})
        img["w1"] = _pack_image(W, 512, 512, 256, 0, (own + 15) // 16, 1, *img.get("w1", (None, None)))
        if self.fight:
            mha = m.att_act
            e = mha.embed_dim
            Wv, bv = mha.in_proj_weight[2 * e:], mha.in_proj_bias[2 * e:]
            Wo, bo = mha.out_proj.weight, mha.out_proj.bias
            Wa = torch.zeros((104, 104), device=self.dev)
            Wa[:100, :100] = (Wo @ Wv).t()
            self.batt.zero_()
            self.batt[:100] = Wo @ bv + bo
            FusedPolicyPair._to_fragments(self.watt, Wa)
            shift, ks, nn = _att_image_geometry(400, 100)
            img["watt"] = _pack_image(Wa, 100, nn, nn, shift, ks, 1, *img.get("watt", (None, None)))
        ws, bs = _lin(m.shared_layer)
        Wsp = torch.zeros((self.NP, self.NW), device=self.dev)
        Wsp[:500, :500] = ws.t()
        FusedPolicyPair._to_fragments(self.ws, Wsp)
        img["ws"] = _pack_image(Wsp, 512, 512, 256, 0, 32, 1, *img.get("ws", (None, None)))
        self.bs.zero_()
        self.bs[:500] = bs
        wa, ba = _lin(m.act_out)
        Wh = torch.zeros((self.NP, self.NH), device=self.dev)
        Wh[:500, :self.n_out] = wa.t()
        FusedPolicyPair._to_fragments(self.wh, Wh)
        img["wh"] = _pack_image(Wh, 32, 32, 32, 0, 32, 8, *img.get("wh", (None, None)))
        self.img = img
        self.bh.zero_()
        self.bh[:self.n_out] = ba

    def fill_chain(self, c, x, n_rows, out=None, act_out=None, rows=None, range_dev=None, act_ptr=None, ld_act=1):
        """Fill one nat.HHPolicyChainEx: x [*, >= d_in] f32 (row stride x.stride(0)); out [*, n_out] f32 and / or
        act_out [*, 4] i32 (per-head argmax); rows (i32) gathers; range_dev (i32 [2], device) = {begin, count}."""
        c.x, c.ldx, c.d_in, c.k1_pad = x.data_ptr(), x.stride(0), self.d_in, self.k1_pad
        c.w1, c.b1 = self.w1.data_ptr(), self.b1.data_ptr()
        if self.fight:
            c.watt, c.batt, c.att_lo, c.att_n, c.att_pad = self.watt.data_ptr(), self.batt.data_ptr(), 400, 100, 104
        else:
            c.watt, c.batt, c.att_lo, c.att_n, c.att_pad = None, None, 0, 0, 0
        c.ws, c.bs, c.wh, c.bh = self.ws.data_ptr(), self.bs.data_ptr(), self.wh.data_ptr(), self.bh.data_ptr()
        c.img_w1, c.us_w1 = (t.data_ptr() for t in self.img["w1"])
        c.img_ws, c.us_ws = (t.data_ptr() for t in self.img["ws"])
        c.img_wh, c.us_wh = (t.data_ptr() for t in self.img["wh"])
        if self.fight:
            c.img_att, c.us_att = (t.data_ptr() for t in self.img["watt"])
        c.out = out.data_ptr() if out is not None else None
        c.ld_out = out.stride(0) if out is not None else 0
        c.act_out = act_ptr if act_ptr is not None else (act_out.data_ptr() if act_out is not None else None)
        c.ld_act = int(ld_act)
        c.rows = rows.data_ptr() if rows is not None else None
        c.range_dev = range_dev.data_ptr() if range_dev is not None else None
        c.n_rows, c.n_out, c.n_heads = int(n_rows), self.n_out, len(self.splits)
        for h in range(4):
            c.head[h] = self.splits[h] if h < len(self.splits) else 0


def run_chains(fill_fns, device, precision: int = 0):
    """One hh_policy_forward_ex launch for up to 8 chains; each fill_fn(chain) fills a nat.HHPolicyChainEx."""
    n = len(fill_fns)
    if n == 0:
        return
    assert n <= 8
    chains = (nat.HHPolicyChainEx * n)()
    for c, f in zip(chains, fill_fns):
        f(c)
    st = torch.cuda.current_stream(device).cuda_stream
    nat.check(nat.lib().hh_policy_forward_ex(n, chains, precision, st), "hh_policy_forward_ex")
