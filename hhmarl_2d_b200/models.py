"""Ray-free restatement of the reference's actor/critic networks (models/ac_models_hetero.py).

Same module tree and parameter names as the reference (`inp1._model.0.weight`, `att_act.in_proj_weight`,
`shared_layer._model.0.weight`, ...), so a state_dict exported from the reference's RLlib models loads
unchanged.  Layer sizes: SURVEY.md A.7 / ac_models_hetero.py:46-83 (Esc1/Esc2), :199-249 (Fight1), :311-361
(Fight2).  `forward(input_dict, state, seq_lens) -> (logits, state)` and `value_function()` keep the ModelV2
calling convention used by env_base.py:392-396; `forward_flat` is the sampler's fast path on the flattened
central observation `[act_1_own | act_2 | obs_1_own | obs_2]` (RLlib flattens Dict spaces in sorted-key order,
train_hetero.py:162-198).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

ACTION_DIM_AC1, ACTION_DIM_AC2 = 4, 3
OBS_AC1, OBS_AC2, OBS_ESC_AC1, OBS_ESC_AC2 = 26, 24, 30, 29
SS_AGENT_AC1, SS_AGENT_AC2 = 12, 10
ACTION_SPLITS = {1: (13, 9, 2, 2), 2: (13, 9, 2)}  # MultiDiscrete heads, env_hetero.py:38-39


class SlimFC(nn.Module):
    """ray.rllib.models.torch.misc.SlimFC: Linear (+ activation), `initializer(weight)`, zero bias."""

    def __init__(self, in_size, out_size, initializer=None, activation_fn=None, use_bias=True, bias_init=0.0):
        super().__init__()
        layers = []
        linear = nn.Linear(in_size, out_size, bias=use_bias)
        if initializer is None:
            initializer = nn.init.xavier_uniform_
        initializer(linear.weight)
        if use_bias:
            nn.init.constant_(linear.bias, bias_init)
        layers.append(linear)
        if activation_fn is not None:
            layers.append(activation_fn())
        self._model = nn.Sequential(*layers)

    def forward(self, x):
        return self._model(x)


def make_shared_layer() -> SlimFC:
    """The module-level SHARED_LAYER singleton of ac_models_hetero.py:22-27: one SlimFC(500, 500, tanh)
    shared by actor and critic paths AND by every policy built in the same process."""
    return SlimFC(500, 500, activation_fn=nn.Tanh, initializer=nn.init.orthogonal_)


def add_time_dimension(x, seq_lens):
    """ray.rllib.policy.rnn_sequencing.add_time_dimension(framework='torch', time_major=False)."""
    b = len(seq_lens) if not torch.is_tensor(seq_lens) else seq_lens.shape[0]
    t = x.shape[0] // b
    return x.reshape(b, t, *x.shape[1:])


def _fc(i, o, act=True):
    return SlimFC(i, o, activation_fn=nn.Tanh if act else None, initializer=nn.init.orthogonal_)


def _attend(mha: nn.MultiheadAttention, x, seq_lens):
    """Self-attention over RLlib's time axis. With one token per sequence softmax over a single key is 1,
    so the block reduces exactly to out_proj(v_proj(x)) -- the sampler's case (seq_lens = [1]*B)."""
    b = x.shape[0]
    n_seq = b if seq_lens is None else (len(seq_lens) if not torch.is_tensor(seq_lens) else seq_lens.shape[0])
    if n_seq == b:
        e = mha.embed_dim
        v = F.linear(x, mha.in_proj_weight[2 * e:], mha.in_proj_bias[2 * e:])
        return mha.out_proj(v)
    xt = x.reshape(n_seq, b // n_seq, x.shape[1])
    out, _ = mha(xt, xt, xt, need_weights=False)
    return out.reshape(b, -1)


class _Base(nn.Module):
    ac_type = 1
    mode = "fight"

    @property
    def own_obs_dim(self):
        return {("fight", 1): OBS_AC1, ("fight", 2): OBS_AC2, ("escape", 1): OBS_ESC_AC1, ("escape", 2): OBS_ESC_AC2}[
            (self.mode, self.ac_type)]

    @property
    def other_obs_dim(self):
        return {("fight", 1): OBS_AC2, ("fight", 2): OBS_AC1, ("escape", 1): OBS_ESC_AC2, ("escape", 2): OBS_ESC_AC1}[
            (self.mode, self.ac_type)]

    @property
    def own_act_dim(self):
        return ACTION_DIM_AC1 if self.ac_type == 1 else ACTION_DIM_AC2

    @property
    def other_act_dim(self):
        return ACTION_DIM_AC2 if self.ac_type == 1 else ACTION_DIM_AC1

    @property
    def central_dim(self):
        return self.own_obs_dim + self.other_obs_dim + 7

    def split_flat(self, flat):
        a, b = self.own_act_dim, self.other_act_dim
        o = self.own_obs_dim
        return {"act_1_own": flat[:, :a], "act_2": flat[:, a:a + b], "obs_1_own": flat[:, a + b:a + b + o],
                "obs_2": flat[:, a + b + o:]}

    def forward_flat(self, flat, seq_lens=None):
        logits, _ = self.forward({"obs": self.split_flat(flat)}, None, seq_lens)
        return logits, self.value_function()

    def get_initial_state(self):
        return [torch.zeros(1)]


class Fight1(_Base):
    """ac_models_hetero.py:181-291 (AC1) / :293-404 (AC2 via subclass)."""
    ac_type = 1
    mode = "fight"

    def __init__(self, shared_layer: SlimFC | None = None, num_outputs: int | None = None):
        super().__init__()
        ss = SS_AGENT_AC1 if self.ac_type == 1 else SS_AGENT_AC2
        own, oth = self.own_obs_dim, self.other_obs_dim
        self.num_outputs = num_outputs if num_outputs is not None else sum(ACTION_SPLITS[self.ac_type])
        self.ss = ss
        self.shared_layer = shared_layer if shared_layer is not None else make_shared_layer()
        self.att_act = nn.MultiheadAttention(100, 2, batch_first=True)
        self.att_val = nn.MultiheadAttention(150, 2, batch_first=True)
        self.inp1 = _fc(ss, 200)
        self.inp2 = _fc(own - ss, 200)
        self.inp3 = _fc(own, 100)
        self.act_out = _fc(500, self.num_outputs, act=False)
        self.v1 = _fc(own + self.own_act_dim, 175)
        self.v2 = _fc(oth + self.other_act_dim, 175)
        self.v3 = _fc(own + self.own_act_dim + oth + self.other_act_dim, 150)
        self.val_out = _fc(500, 1, act=False)
        self._val = None

    def forward(self, input_dict, state=None, seq_lens=None):
        o = input_dict["obs"]
        own = o["obs_1_own"]
        v1_in = torch.cat((own, o["act_1_own"]), dim=1)
        v2_in = torch.cat((o["obs_2"], o["act_2"]), dim=1)
        v3_in = torch.cat((v1_in, v2_in), dim=1)
        x = torch.cat((self.inp1(own[:, :self.ss]), self.inp2(own[:, self.ss:])), dim=1)
        x_full = self.inp3(own)
        x_full = F.normalize(x_full + _attend(self.att_act, x_full, seq_lens))
        x = self.act_out(self.shared_layer(torch.cat((x, x_full), dim=1)))
        y = torch.cat((self.v1(v1_in), self.v2(v2_in)), dim=1)
        y_full = self.v3(v3_in)
        y_full = F.normalize(y_full + _attend(self.att_val, y_full, seq_lens))
        self._val = self.val_out(self.shared_layer(torch.cat((y, y_full), dim=1)))
        return x, []

    def value_function(self):
        assert self._val is not None, "must call forward first!"
        return torch.reshape(self._val, [-1])

    def actor(self, own, seq_lens=None):
        """Actor head only (what a frozen opponent needs: the critic inputs are all zeros, env_base.py:357-371)."""
        x = torch.cat((self.inp1(own[:, :self.ss]), self.inp2(own[:, self.ss:])), dim=1)
        x_full = self.inp3(own)
        x_full = F.normalize(x_full + _attend(self.att_act, x_full, seq_lens))
        return self.act_out(self.shared_layer(torch.cat((x, x_full), dim=1)))


class Fight2(Fight1):
    ac_type = 2


class Esc1(_Base):
    """ac_models_hetero.py:29-103 (AC1) / :105-179 (AC2 via subclass)."""
    ac_type = 1
    mode = "escape"

    def __init__(self, shared_layer: SlimFC | None = None, num_outputs: int | None = None):
        super().__init__()
        self.num_outputs = num_outputs if num_outputs is not None else sum(ACTION_SPLITS[self.ac_type])
        self.k1 = 7 if self.ac_type == 1 else 6
        self.shared_layer = shared_layer if shared_layer is not None else make_shared_layer()
        self.inp1 = _fc(self.k1, 150)
        self.inp2 = _fc(18, 250)
        self.inp3 = _fc(5, 100)
        self.act_out = _fc(500, self.num_outputs, act=False)
        self.inp1_val = _fc(OBS_ESC_AC1 + ACTION_DIM_AC1 + OBS_ESC_AC2 + ACTION_DIM_AC2, 500)
        self.val_out = _fc(500, 1, act=False)
        self._v1 = None

    def forward(self, input_dict, state=None, seq_lens=None):
        o = input_dict["obs"]
        own = o["obs_1_own"]
        k = self.k1
        self._v1 = torch.cat((own, o["act_1_own"], o["obs_2"], o["act_2"]), dim=1)
        x = torch.cat((self.inp1(own[:, :k]), self.inp2(own[:, k:k + 18]), self.inp3(own[:, k + 18:])), dim=1)
        return self.act_out(self.shared_layer(x)), []

    def value_function(self):
        assert self._v1 is not None, "must call forward first!"
        return torch.reshape(self.val_out(self.shared_layer(self.inp1_val(self._v1))), [-1])

    def actor(self, own, seq_lens=None):
        k = self.k1
        x = torch.cat((self.inp1(own[:, :k]), self.inp2(own[:, k:k + 18]), self.inp3(own[:, k + 18:])), dim=1)
        return self.act_out(self.shared_layer(x))


class Esc2(Esc1):
    ac_type = 2


class CommanderGru(nn.Module):
    """models/ac_models_hier.py:21-112: commander policy for the 3-vs-3 HighLevelEnv.  Actor: 4/20/10-wide
    branches + GRU(200) over the RLlib time axis on the full 34-d observation; centralised critic: own and the
    two team-mates' (observation, action) pairs + GRU(200); both through the shared 500x500 layer.
    State = [actor GRU h (200), critic GRU h (200)]."""
    OBS_DIM, N_OPP_HL, OBS_OPP = 34, 2, 10

    def __init__(self, shared_layer: SlimFC | None = None, num_outputs: int = 3):
        super().__init__()
        self.num_outputs = num_outputs
        self.shared_layer = shared_layer if shared_layer is not None else make_shared_layer()
        self.rnn_act = nn.GRU(200, 200, batch_first=True)
        self.rnn_val = nn.GRU(200, 200, batch_first=True)
        self.inp1 = _fc(4, 50)
        self.inp2 = _fc(self.N_OPP_HL * self.OBS_OPP, 200)
        self.inp3 = _fc(10, 50)
        self.inp4 = _fc(self.OBS_DIM, 200)
        self.act_out = _fc(500, num_outputs, act=False)
        self.v1 = _fc(self.OBS_DIM + 1, 100)
        self.v2 = _fc(self.OBS_DIM + 1, 100)
        self.v3 = _fc(self.OBS_DIM + 1, 100)
        self.v4 = _fc(3 * (self.OBS_DIM + 1), 200)
        self.val_out = _fc(500, 1, act=False)
        self._val = None

    def get_initial_state(self):
        return [torch.zeros(200), torch.zeros(200)]

    def forward(self, input_dict, state, seq_lens):
        o = input_dict["obs"]
        own = o["obs_1_own"]
        k = 4 + self.N_OPP_HL * self.OBS_OPP
        v1 = torch.cat((own, o["act_1_own"]), dim=1)
        v2 = torch.cat((o["obs_2"], o["act_2"]), dim=1)
        v3 = torch.cat((o["obs_3"], o["act_3"]), dim=1)
        v4 = torch.cat((v1, v2, v3), dim=1)
        x = torch.cat((self.inp1(own[:, :4]), self.inp2(own[:, 4:k]), self.inp3(own[:, k:])), dim=1)
        x_full = self.inp4(own)
        y, h = self.rnn_act(add_time_dimension(x_full, seq_lens), torch.unsqueeze(state[0], 0))
        x_full = F.normalize(x_full + y.reshape(-1, 200))
        x = self.act_out(self.shared_layer(torch.cat((x, x_full), dim=1)))
        z = torch.cat((self.v1(v1), self.v2(v2), self.v3(v3)), dim=1)
        z_full = self.v4(v4)
        w, kk = self.rnn_val(add_time_dimension(z_full, seq_lens), torch.unsqueeze(state[1], 0))
        z_full = F.normalize(z_full + w.reshape(-1, 200))
        self._val = self.val_out(self.shared_layer(torch.cat((z, z_full), dim=1)))
        return torch.reshape(x, [-1, self.num_outputs]), [torch.squeeze(h, 0), torch.squeeze(kk, 0)]

    def value_function(self):
        assert self._val is not None, "must call forward first!"
        return torch.reshape(self._val, [-1])


def build_policy_pair(mode: str = "fight", shared_layer: SlimFC | None = None):
    """(ac1_policy model, ac2_policy model) sharing one SHARED_LAYER, like the reference process does."""
    shared = shared_layer if shared_layer is not None else make_shared_layer()
    if mode == "fight":
        return Fight1(shared), Fight2(shared)
    return Esc1(shared), Esc2(shared)


def deterministic_actions(logits: torch.Tensor, ac_type: int) -> torch.Tensor:
    """Per-head argmax (env_base.py:373-382): int64 [B, 4|3]."""
    outs, o = [], 0
    for n in ACTION_SPLITS[ac_type]:
        outs.append(torch.argmax(logits[:, o:o + n], dim=-1))
        o += n
    return torch.stack(outs, dim=1)


def fill_from_seed(model: nn.Module, seed: int, scale: float = 0.2):
    """Deterministic, platform-independent weights for parity fixtures (numpy PCG64, NOT torch's RNG)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    seen = set()
    with torch.no_grad():
        for name, p in sorted(model.named_parameters(), key=lambda kv: kv[0]):
            if id(p) in seen:
                continue
            seen.add(id(p))
            w = rng.standard_normal(p.numel()).astype(np.float32) * (scale / max(1.0, p.shape[-1] ** 0.5) if p.dim() > 1 else 0.05)
            p.copy_(torch.from_numpy(w).reshape(p.shape))
    return model
