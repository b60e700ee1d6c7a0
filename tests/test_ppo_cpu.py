"""Host-side logic of the learner / sampler that does not need a GPU: the sequence re-ordering used for
RLlib-style max_seq_len chunks, the MultiCategorical helpers, a CPU PPO update, and the 2-rank gloo
gradient exchange (world_size 2, as the multi-GPU path uses NCCL)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hhmarl_2d_b200 import models as M
from hhmarl_2d_b200.ppo import PPOLearner, _seq_major
from hhmarl_2d_b200.sampler import multicategorical_logp_entropy_kl, multicategorical_sample


def _fake_batch(T, N, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    acts = torch.stack([torch.randint(0, 13, (T, N, 2), generator=g), torch.randint(0, 9, (T, N, 2), generator=g),
                        torch.randint(0, 2, (T, N, 2), generator=g), torch.randint(0, 2, (T, N, 2), generator=g)], -1).int()
    return dict(flat1=r(T, N, 57), flat2=r(T, N, 57), actions=acts, logits1=r(T, N, 26), logits2=r(T, N, 24),
                logp=-r(T, N, 2) - 3, vf=r(T, N, 2), rew=r(T, N, 2), adv=r(T, N, 2) - 0.5, vtarg=r(T, N, 2),
                done=torch.zeros(T, N, dtype=torch.uint8))


def test_seq_major_keeps_sequences_contiguous():
    T, N, L = 40, 3, 20
    x = torch.arange(T)[:, None].repeat(1, N) * 100 + torch.arange(N)[None, :]        # value = t*100 + n
    y = _seq_major(x, L).reshape(T // L, N, L)
    for c in range(T // L):
        for n in range(N):
            assert (y[c, n] == torch.arange(c * L, (c + 1) * L) * 100 + n).all()


def test_multicategorical_helpers():
    torch.manual_seed(0)
    logits = torch.randn(4096, 26)
    a, logp = multicategorical_sample(logits, (13, 9, 2, 2))
    lp2, ent, kl = multicategorical_logp_entropy_kl(logits, a, (13, 9, 2, 2), logits)
    assert torch.allclose(logp, lp2, atol=1e-6) and kl.abs().max() < 1e-6 and (ent > 0).all()
    a_det, _ = multicategorical_sample(logits, (13, 9, 2, 2), explore=False)
    assert (a_det == M.deterministic_actions(logits, 1)).all()
    # sampling frequencies follow softmax (first head of the first row repeated)
    rep = logits[:1].repeat(20000, 1)
    s, _ = multicategorical_sample(rep, (13, 9, 2, 2))
    freq = torch.bincount(s[:, 0], minlength=13).float() / 20000
    assert (freq - torch.softmax(logits[0, :13], 0)).abs().max() < 0.02


def test_ppo_update_runs_and_improves_surrogate_on_cpu():
    torch.manual_seed(0)
    m1, m2 = M.build_policy_pair("fight")
    learner = PPOLearner(m1, m2, num_sgd_iter=2, sgd_minibatch_size=80, lr=1e-3)
    batch = _fake_batch(20, 16)
    with torch.no_grad():       # make old logits / logp consistent with the current policies
        for i, (m, k) in enumerate(((m1, "1"), (m2, "2"))):
            lg, vf = m.forward_flat(batch["flat" + k].reshape(-1, 57))
            batch["logits" + k] = lg.reshape(20, 16, -1)
            n = 4 if i == 0 else 3
            lp, _, _ = multicategorical_logp_entropy_kl(lg, batch["actions"][:, :, i, :n].reshape(-1, n), learner.splits[i])
            batch["logp"][:, :, i] = lp.reshape(20, 16)
    w0 = m1.shared_layer._model[0].weight.clone()
    st = learner.update(batch)
    assert st["minibatches"] == 2 * (16 // 4) and np.isfinite(st["loss"])
    assert not torch.equal(w0, m1.shared_layer._model[0].weight)          # the shared layer trains
    assert m1.shared_layer is m2.shared_layer


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m1, m2 = M.build_policy_pair("fight")
    learner = PPOLearner(m1, m2, num_sgd_iter=1, sgd_minibatch_size=80)
    batch = _fake_batch(20, 8, seed=rank)                                  # different shards per rank
    learner.update(batch)
    q.put((rank, float(sum(p.sum() for p in learner.params))))
    dist.destroy_process_group()


def test_gradient_allreduce_keeps_two_ranks_in_sync():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert abs(res[0] - res[1]) < 1e-4 * max(1.0, abs(res[0]))             # identical weights after the update


def test_policy_files_round_trip_with_reference_naming(tmp_path):
    from hhmarl_2d_b200 import checkpoint
    d = str(tmp_path / "policies")
    saved = {}
    for level in (3, 4):
        f1, f2 = M.build_policy_pair("fight")
        M.fill_from_seed(f1, level); M.fill_from_seed(f2, level + 10)     # (the pair shares SHARED_LAYER)
        checkpoint.save_policies(d, level, "fight", f1, f2)
        saved[level] = f1
    e1, e2 = M.build_policy_pair("escape")
    checkpoint.save_policies(d, 3, "escape", e1, e2)
    assert sorted(os.listdir(d)) == ["L3_AC1_escape.pt", "L3_AC1_fight.pt", "L3_AC2_escape.pt", "L3_AC2_fight.pt",
                                     "L4_AC1_fight.pt", "L4_AC2_fight.pt"]
    l4 = checkpoint.load_opponent_policies(d, 4, "fight")
    assert set(l4) == {"fight_1", "fight_2"}
    l5 = checkpoint.load_opponent_policies(d, 5, "fight")
    assert set(l5) == {3, 4, 5} and set(l5[5]) == {"escape_1", "escape_2"}
    x = torch.rand(4, 26)
    with torch.no_grad():
        assert torch.equal(l5[3]["fight_1"].actor(x), saved[3].actor(x))
        assert torch.equal(l5[4]["fight_1"].actor(x), saved[4].actor(x))
    # HighLevelEnv container (env_base.py:332-346): fight policies of eval_level_ag, escape L5 else L3, "_opp" pair
    # only for the low-level evaluation mode
    hl = checkpoint.load_highlevel_policies(d, eval_level_ag=4)
    assert set(hl) == {"fight_1", "fight_2", "escape_1", "escape_2"}
    hl = checkpoint.load_highlevel_policies(d, eval_level_ag=4, eval_level_opp=3, eval_hl=False)
    assert set(hl) == {"fight_1", "fight_2", "escape_1", "escape_2", "fight_1_opp", "fight_2_opp"}
    xe = torch.rand(4, 30)
    with torch.no_grad():
        assert torch.equal(hl["fight_1"].actor(x), saved[4].actor(x)) and torch.equal(hl["fight_1_opp"].actor(x), saved[3].actor(x))
        assert torch.equal(hl["escape_1"].actor(xe), e1.actor(xe))
    e51, e52 = M.build_policy_pair("escape")
    M.fill_from_seed(e51, 77); M.fill_from_seed(e52, 78)
    checkpoint.save_policies(d, 5, "escape", e51, e52)
    with torch.no_grad():
        assert torch.equal(checkpoint.load_highlevel_policies(d, 4)["escape_1"].actor(xe), e51.actor(xe))


def test_training_state_resumes_and_small_batches_still_learn(tmp_path):
    """checkpoint.save_training_state / load_training_state (what algo.save() keeps beyond the exported weights: Adam
    moments, adaptive KL coefficients, epoch, shuffling RNG): a resumed learner continues bit-identically.  Also: a
    batch smaller than sgd_minibatch_size is ONE minibatch (not zero), the remainder minibatch is used, and two
    policies built separately (two shared layers) are refused by the packed forward."""
    from hhmarl_2d_b200 import checkpoint
    from hhmarl_2d_b200.fused_forward import PackedPolicyPair
    torch.manual_seed(0)

    def fresh():
        torch.manual_seed(1)
        m1, m2 = M.build_policy_pair("fight")
        return m1, m2, PPOLearner(m1, m2, num_sgd_iter=1, sgd_minibatch_size=8192, lr=1e-3)

    b1, b2 = _fake_batch(20, 5, seed=1), _fake_batch(20, 5, seed=2)
    m1, m2, la = fresh()
    st = la.update(b1)
    assert st["minibatches"] == 1                      # 5 sequences < 8192 / 20: one (smaller) minibatch, weights move
    checkpoint.save_training_state(str(tmp_path), 3, "fight", la)
    la.update(b2)
    want = la.flat.detach().clone()
    n1, n2, lb = fresh()
    assert checkpoint.load_training_state(str(tmp_path), 3, "fight", lb) == 1
    assert torch.equal(n1.act_out._model[0].weight, lb.params[[id(p) for p in lb.params].index(id(n1.act_out._model[0].weight))])
    lb.update(b2)
    assert torch.equal(lb.flat.detach(), want) and lb.kl_coeff == la.kl_coeff and lb.epoch == 2
    # remainder minibatch: 5 sequences, 2 per minibatch -> 3 minibatches
    _, _, lc = fresh()
    lc.mb = 40
    assert lc.update(b1)["minibatches"] == 3
    # separately built policies do not share a shared layer
    with pytest.raises(ValueError):
        PackedPolicyPair(M.Fight1(), M.Fight2())
    # policy files of one level must agree on the shared layer
    checkpoint.save_policies(str(tmp_path), 3, "fight", m1, m2)
    checkpoint.load_pair(str(tmp_path), 3, "fight")
    o1, o2 = M.build_policy_pair("fight")
    with torch.no_grad():
        o2.shared_layer._model[0].weight.add_(1.0)
    blob = torch.load(checkpoint.policy_path(str(tmp_path), 3, 2, "fight"))
    blob["state_dict"] = o2.state_dict()
    torch.save(blob, checkpoint.policy_path(str(tmp_path), 3, 2, "fight"))
    with pytest.raises(ValueError):
        checkpoint.load_pair(str(tmp_path), 3, "fight")
