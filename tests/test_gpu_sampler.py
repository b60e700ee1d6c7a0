"""On-device sampler / GAE / PPO on the GPU: buffer layout (central-critic observation, action write-back),
consistency of the recorded trajectory with an independent replay of the env, the GAE kernel against a numpy
restatement of RLlib's compute_advantages, and one PPO update."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gae_numpy(rew, vf, last_vf, done, gamma, lam):
    T = rew.shape[0]
    adv = np.zeros_like(rew)
    nxt, gae = last_vf.copy(), np.zeros_like(last_vf)
    for t in range(T - 1, -1, -1):
        nt = (1.0 - done[t].astype(np.float32))[:, None]
        delta = rew[t] + gamma * nxt * nt - vf[t]
        gae = delta + gamma * lam * nt * gae
        adv[t] = gae
        nxt = vf[t]
    return adv, adv + vf


def test_gae_kernel_matches_numpy():
    from hhmarl_2d_b200 import _native as nat
    rng = np.random.default_rng(0)
    T, N = 37, 1000
    rew = rng.normal(size=(T, N, 2)).astype(np.float32); vf = rng.normal(size=(T, N, 2)).astype(np.float32)
    last = rng.normal(size=(N, 2)).astype(np.float32); done = (rng.random((T, N)) < 0.05).astype(np.uint8)
    d = lambda x: torch.from_numpy(x).cuda()
    adv, vt = torch.empty(T, N, 2, device="cuda"), torch.empty(T, N, 2, device="cuda")
    r_, v_, l_, d_ = d(rew), d(vf), d(last), d(done)
    nat.check(nat.lib().hh_gae(T, N, r_.data_ptr(), v_.data_ptr(), l_.data_ptr(), d_.data_ptr(), 0.99, 0.95,
                               adv.data_ptr(), vt.data_ptr(), torch.cuda.current_stream().cuda_stream), "hh_gae")
    ea, ev = _gae_numpy(rew, vf, last, done, 0.99, 0.95)
    np.testing.assert_allclose(adv.cpu().numpy(), ea, rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(vt.cpu().numpy(), ev, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("use_graph,mode,level", [(False, "fight", 3), (True, "fight", 3), (True, "escape", 3), (False, "escape", 5),
                                                  (True, "fight", 5)])
def test_sampler_fragment_is_consistent(use_graph, mode, level):
    """Fight (26 / 24 observations) and escape (30 / 29; env_base.py:137-164, 223-233) rollouts, scripted (3) and frozen-policy
    (5) opponents: buffers, central-observation layout, action write-back, log-probs and GAE of recorded fragments."""
    from hhmarl_2d_b200 import VecLowLevelEnv, make_args, VecSampler, TorchPolicy
    from hhmarl_2d_b200 import models as M
    from hhmarl_2d_b200.sampler import multicategorical_logp_entropy_kl
    torch.manual_seed(0)
    n, T = 512, 40
    mk = lambda: VecLowLevelEnv(n, make_args(level=level, agent_mode=mode), device=0, seed=11, allow_standin_opponents=True)  # noqa: E731
    env = mk()
    d1, d2 = env.obs_dim
    assert (d1, d2) == ((26, 24) if mode == "fight" else (30, 29))
    m1, m2 = M.build_policy_pair(mode)
    m1.cuda(); m2.cuda()
    smp = VecSampler(env, TorchPolicy(m1, 1), TorchPolicy(m2, 2), fragment_len=T, use_cuda_graph=use_graph)
    assert smp.direct and smp.native_glue
    frags = [{k: v.clone() for k, v in smp.collect().items()} for _ in range(3)]
    # independent replay of the recorded actions on a fresh env with the same seed
    env2 = mk()
    o1, o2 = env2.reset()
    for b in frags:
        for t in range(T):
            # observation columns of the flattened central obs (train_hetero.py:162-181, sorted-key order)
            assert torch.equal(b["flat1"][t][:, 7:7 + d1], o1) and torch.equal(b["flat1"][t][:, 7 + d1:], o2)
            assert torch.equal(b["flat2"][t][:, 7:7 + d2], o2) and torch.equal(b["flat2"][t][:, 7 + d2:], o1)
            o1, o2, r, d = env2.step(b["actions"][t].contiguous())
            assert torch.equal(r, b["rew"][t]) and torch.equal(d, b["done"][t])
        # action write-back of on_postprocess_trajectory (train_hetero.py:140-160)
        a = b["actions"].float()
        assert torch.allclose(b["flat1"][:, :, 0], a[:, :, 0, 0] / 12) and torch.allclose(b["flat1"][:, :, 1], a[:, :, 0, 1] / 8)
        assert torch.equal(b["flat1"][:, :, 2:4], a[:, :, 0, 2:4]) and torch.allclose(b["flat1"][:, :, 4], a[:, :, 1, 0] / 12)
        assert torch.allclose(b["flat2"][:, :, 0], a[:, :, 1, 0] / 12) and torch.allclose(b["flat2"][:, :, 3], a[:, :, 0, 0] / 12)
        assert torch.equal(b["flat2"][:, :, 6], a[:, :, 0, 3])
        # actions inside the MultiDiscrete ranges, log-probs consistent with the recorded logits
        assert (b["actions"][..., 0] < 13).all() and (b["actions"][..., 1] < 9).all() and (b["actions"][..., 2:] < 2).all()
        lp, _, _ = multicategorical_logp_entropy_kl(b["logits1"].reshape(-1, 26), b["actions"][:, :, 0, :].reshape(-1, 4), (13, 9, 2, 2))
        assert torch.allclose(lp.reshape(T, n), b["logp"][:, :, 0], atol=1e-5)
        # GAE of the fragment against the numpy restatement
        ea, ev = _gae_numpy(b["rew"].cpu().numpy(), b["vf"].cpu().numpy(), b["last_vf"].cpu().numpy(), b["done"].cpu().numpy(), 0.99, 0.95)
        np.testing.assert_allclose(b["adv"].cpu().numpy(), ea, rtol=1e-3, atol=1e-3)
    assert sum(int(b["done"].sum()) for b in frags) > 0


def test_compute_actions_contract_on_the_fused_path():
    """TorchPolicy.compute_actions = RLlib's Policy.compute_actions (train_hetero.py:200-205, 242): (actions, [], {action_logp,
    action_dist_inputs, vf_preds}); once a sampler exists it runs the SAME fused tcgen05 forward as the rollout, checked
    here against the eager torch modules."""
    from hhmarl_2d_b200 import VecLowLevelEnv, make_args, VecSampler, TorchPolicy
    from hhmarl_2d_b200 import models as M
    from hhmarl_2d_b200.sampler import multicategorical_logp_entropy_kl
    torch.manual_seed(2)
    for mode in ("fight", "escape"):
        m1, m2 = M.build_policy_pair(mode)
        m1.cuda(); m2.cuda()
        p1, p2 = TorchPolicy(m1, 1), TorchPolicy(m2, 2)
        eager = p1.compute_actions(torch.rand(7, m1.central_dim, device="cuda"))           # before attach: torch forward
        assert eager[0].shape == (7, 4) and eager[1] == []
        env = VecLowLevelEnv(64, make_args(level=3, agent_mode=mode), device=0, seed=1)
        VecSampler(env, p1, p2, fragment_len=20, use_cuda_graph=False)
        assert p1._fused is not None and p2._fused is not None
        for pol, m, heads in ((p1, m1, (13, 9, 2, 2)), (p2, m2, (13, 9, 2))):
            B = 777
            obs = torch.rand(B, m.central_dim, device="cuda")
            obs[:, :7] = 0
            act, state, extra = pol.compute_actions(obs, explore=True)
            assert state == [] and set(extra) == {"action_logp", "action_dist_inputs", "vf_preds"}
            assert act.shape == (B, len(heads)) and extra["action_dist_inputs"].shape == (B, sum(heads))
            assert extra["action_logp"].shape == (B,) and extra["vf_preds"].shape == (B,)
            with torch.no_grad():
                lg, vf = m.forward_flat(obs)
            assert (extra["action_dist_inputs"] - lg).abs().max().item() < 3e-5 and (extra["vf_preds"] - vf).abs().max().item() < 3e-5
            lp, _, _ = multicategorical_logp_entropy_kl(extra["action_dist_inputs"], act, heads)
            assert torch.allclose(lp, extra["action_logp"], atol=1e-5)
            for h, w in enumerate(heads):
                assert int(act[:, h].min()) >= 0 and int(act[:, h].max()) < w
            det, _, x2 = pol.compute_actions(obs, explore=False)
            assert torch.equal(det, M.deterministic_actions(x2["action_dist_inputs"], pol.ac_type))
            a1, _, x1 = pol.compute_single_action(obs[3], explore=False)
            assert torch.equal(a1, det[3]) and x1["vf_preds"].shape == ()


@pytest.mark.parametrize("level,mode", [(3, "escape"), (5, "escape")])
def test_escape_mode_training_iterations(level, mode):
    """the escape-mode path that PRODUCES the escape policies levels 5 / hier consume -- sample with the fused
    forward (30 / 29 observations, ammunition penalties env_base.py:223-233) and learn for three PPO iterations."""
    from hhmarl_2d_b200 import VecLowLevelEnv, make_args, VecSampler, TorchPolicy, PPOLearner
    from hhmarl_2d_b200 import models as M
    torch.manual_seed(0)
    env = VecLowLevelEnv(256, make_args(level=level, agent_mode=mode, esc_dist_rew=True), device=0, seed=5, allow_standin_opponents=True)
    m1, m2 = M.build_policy_pair(mode)
    m1.cuda(); m2.cuda()
    smp = VecSampler(env, TorchPolicy(m1, 1), TorchPolicy(m2, 2), fragment_len=20, use_cuda_graph=True)
    learner = PPOLearner(m1, m2, num_sgd_iter=2, sgd_minibatch_size=2560)
    w0 = m1.inp2._model[0].weight.clone()     # (a live autograd graph of a parameter on the default stream: must not break the capture)
    for it in range(3):
        b = smp.collect()
        assert b["flat1"].shape[-1] == 7 + 30 + 29 and torch.isfinite(b["adv"]).all()
        st = learner.update(b)
        smp.refresh_policy()
        assert st["minibatches"] == 4 and np.isfinite(st["loss"])
    assert not torch.equal(w0, m1.inp2._model[0].weight)
    with torch.no_grad():      # the sampler's packed weights follow the learner (refresh_policy)
        b = smp.collect()
        f1 = b["flat1"][3].clone(); f1[:, :7] = 0
        lg, _ = m1.forward_flat(f1)
    assert (b["logits1"][3] - lg).abs().max().item() < 5e-5


def test_learner_improves_the_level1_reward():
    """Wiring check of sampler -> GAE -> action write-back -> learner (train_hetero.py:120-160, 212-217): at level 1 (static
    opponents) the mean fragment reward of a PPO-trained pair must rise clearly above that of the initial random policy."""
    from hhmarl_2d_b200 import VecLowLevelEnv, make_args, VecSampler, TorchPolicy, PPOLearner
    from hhmarl_2d_b200 import models as M
    torch.manual_seed(0)
    n, T = 2048, 40
    env = VecLowLevelEnv(n, make_args(level=1), device=0, seed=21)
    m1, m2 = M.build_policy_pair("fight")
    m1.cuda(); m2.cuda()
    smp = VecSampler(env, TorchPolicy(m1, 1), TorchPolicy(m2, 2), fragment_len=T, use_cuda_graph=True)
    learner = PPOLearner(m1, m2, lr=3e-4, num_sgd_iter=4, sgd_minibatch_size=8192)
    hist = []
    for it in range(30):
        b = smp.collect()
        hist.append(float(b["rew"].sum(0).mean()))          # mean reward per arena and fragment, both agents
        learner.update(b)
        smp.refresh_policy()
    first, last = np.mean(hist[:4]), np.mean(hist[-4:])
    print("level-1 fragment reward: first 4 iterations %.3f, last 4 %.3f" % (first, last), hist)
    assert last > first + 0.15, (first, last)


@pytest.mark.parametrize("use_graph", [False, True])
def test_groups_on_their_own_streams_collect_the_same_fragment(use_graph):
    """VecSampler(groups=g) advances g equal parts of the batch on their own streams (hh_step_range; the kernels of the whole batch
    pointed at each part's rows): every buffer of the fragment must equal the one-stream sampler's bit for bit, twice in a row."""
    from hhmarl_2d_b200 import VecLowLevelEnv, make_args, VecSampler, TorchPolicy
    from hhmarl_2d_b200 import models as M
    torch.manual_seed(0)
    m1, m2 = M.build_policy_pair("fight")
    m1.cuda(); m2.cuda()
    n, out = 1024, []
    for groups in (1, 2, 8):
        env = VecLowLevelEnv(n, make_args(level=3), device=0, seed=11, autoreset=True)
        smp = VecSampler(env, TorchPolicy(m1, 1), TorchPolicy(m2, 2), fragment_len=20, use_cuda_graph=use_graph, groups=groups)
        assert smp.groups == groups
        frs = []
        for _ in range(3 if use_graph else 2):           # (with a graph: warm-up / capture, then two replays)
            frs.append({k: v.clone() for k, v in smp.collect().items()})
        out.append(frs)
    for other in out[1:]:
        for fa, fb in zip(out[0], other):
            for k in fa:
                assert torch.equal(fa[k], fb[k]), k
    # hh_fragment_writeback = on_postprocess_trajectory (train_hetero.py:120-160): scaled actions in the critics' action columns
    f = out[0][-1]
    a = f["actions"].float()
    sc = torch.tensor([12.0, 8.0, 1.0, 1.0], device="cuda")
    own1, own2 = a[:, :, 0, :] / sc, a[:, :, 1, :3] / sc[:3]
    assert torch.equal(f["flat1"][:, :, 0:4], own1) and torch.equal(f["flat1"][:, :, 4:7], own2)
    assert torch.equal(f["flat2"][:, :, 0:3], own2) and torch.equal(f["flat2"][:, :, 3:7], own1)
    assert VecSampler(VecLowLevelEnv(4096, make_args(level=3), device=0, seed=1), TorchPolicy(m1, 1), TorchPolicy(m2, 2)).groups == 4
    with pytest.raises(ValueError):
        VecSampler(VecLowLevelEnv(640, make_args(level=3), device=0, seed=1), TorchPolicy(m1, 1), TorchPolicy(m2, 2), groups=2)


def test_collect_host_with_two_buffer_sets_delivers_every_fragment():
    """VecSampler(n_buffers=2).collect_host(): fragment k is copied to pinned host memory on a copy stream while fragment k + 1 is
    sampled into the other buffer set; the host dicts must equal what a single-buffered sampler with the same seed collects."""
    from hhmarl_2d_b200 import VecLowLevelEnv, make_args, VecSampler, TorchPolicy
    from hhmarl_2d_b200 import models as M
    torch.manual_seed(0)
    m1, m2 = M.build_policy_pair("fight")
    m1.cuda(); m2.cuda()
    n = 1024
    ref = VecSampler(VecLowLevelEnv(n, make_args(level=3), device=0, seed=7, autoreset=True), TorchPolicy(m1, 1), TorchPolicy(m2, 2),
                     fragment_len=20)
    dbl = VecSampler(VecLowLevelEnv(n, make_args(level=3), device=0, seed=7, autoreset=True), TorchPolicy(m1, 1), TorchPolicy(m2, 2),
                     fragment_len=20, n_buffers=2)
    want = [{k: v.cpu() for k, v in ref.collect().items()} for _ in range(5)]
    pending = []
    for f in range(5):
        h, ev = dbl.collect_host()                       # no wait here: the next call overlaps this copy
        pending.append((f, h, ev))
        if len(pending) == 2:                            # read fragment f - 1 before the call after next reuses its buffers
            g, hh, e = pending.pop(0)
            e.synchronize()
            for k in want[g]:
                assert torch.equal(hh[k], want[g][k]), (g, k)
    g, hh, e = pending.pop(0)
    e.synchronize()
    for k in want[g]:
        assert torch.equal(hh[k], want[g][k]), (g, k)
    assert all(v.is_pinned() for v in hh.values())


def test_fused_multicategorical_terms_match_torch():
    """hh_multicat_forward / _backward (the learner's logp / entropy / KL of a MultiDiscrete policy in one kernel each) against the
    per-head torch expression: values and the gradient with respect to the logits, both policies' head layouts, strided inputs."""
    from hhmarl_2d_b200.sampler import multicategorical_logp_entropy_kl as f
    torch.manual_seed(0)
    for splits in ((13, 9, 2, 2), (13, 9, 2)):
        n, w = 3000, sum(splits)
        z = (3 * torch.randn(n, w, device="cuda")).requires_grad_()
        zo = (3 * torch.randn(n, w + 5, device="cuda"))[:, :w]                     # row stride != width
        act = torch.stack([torch.randint(0, k, (n,), device="cuda") for k in splits] + [torch.zeros(n, device="cuda", dtype=torch.long)] * (4 - len(splits)),
                          dim=1).to(torch.int32)[:, :len(splits)]                    # int32, row stride 4
        cw = torch.randn(3, n, device="cuda")
        lp, en, kl = f(z, act, splits, zo)                                          # fused (CUDA, int32 actions, old logits given)
        (gz,) = torch.autograd.grad((cw[0] * lp + cw[1] * en + cw[2] * kl).sum(), z)
        z2 = z.detach().clone().requires_grad_()
        lp2, en2, kl2 = f(z2, act.long(), splits, zo)                               # int64 actions -> the per-head torch path
        (gz2,) = torch.autograd.grad((cw[0] * lp2 + cw[1] * en2 + cw[2] * kl2).sum(), z2)
        for a, b, name in ((lp, lp2, "logp"), (en, en2, "entropy"), (kl, kl2, "kl"), (gz, gz2, "grad")):
            err = (a - b).abs().max().item()
            assert err <= 2e-5 * max(1.0, b.abs().max().item()), (splits, name, err)


def test_fused_ppo_loss_matches_the_torch_expression():
    """hh_ppo_loss (clipped surrogate + KL + clamped value loss + entropy bonus of one policy in one kernel) against the torch
    expression PPOLearner uses on the CPU: the four means and the gradients with respect to logp / entropy / kl / vf, with ratios
    inside and outside the clip range, both advantage signs, value errors beyond vf_clip and a non-zero entropy coefficient."""
    from hhmarl_2d_b200.ppo import _FusedPPOLoss
    torch.manual_seed(1)
    n, clip, vf_clip, vf_coeff, ent_coeff = 5000, 0.25, 10.0, 1.0, 0.01
    klc = torch.tensor([0.3], device="cuda")
    old_logp, adv = torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
    vtarg = 3 * torch.randn(n, device="cuda")
    base = [old_logp + 0.4 * torch.randn(n, device="cuda"), torch.rand(n, device="cuda"), torch.rand(n, device="cuda"),
            vtarg + 4 * torch.randn(n, device="cuda")]
    outs = []
    for fused in (True, False):
        logp, ent, kl, vf = (t.clone().requires_grad_() for t in base)
        if fused:
            o = _FusedPPOLoss.apply(logp, ent, kl, vf, old_logp, adv, vtarg, klc, clip, vf_clip, vf_coeff, ent_coeff)
        else:
            ratio = torch.exp(logp - old_logp)
            surr = torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - clip, 1 + clip))
            vfl = torch.clamp((vf - vtarg) ** 2, 0, vf_clip)
            o = torch.stack([(-surr + klc[0] * kl + vf_coeff * vfl - ent_coeff * ent).mean(), kl.mean(), vfl.mean(), ent.mean()])
        g = torch.autograd.grad(o[0] * 1.7, (logp, ent, kl, vf))
        outs.append((o.detach(), g))
    (of, gf), (ot, gt) = outs
    assert (of - ot).abs().max().item() <= 1e-5 * max(1.0, ot.abs().max().item()), (of, ot)
    for a, b, name in zip(gf, gt, ("logp", "entropy", "kl", "vf")):
        assert (a - b).abs().max().item() <= 1e-6 * max(1.0, b.abs().max().item()) + 1e-9, name
    assert ((torch.exp(base[0] - old_logp) - 1).abs() > clip).float().mean().item() > 0.2      # the clip branches are exercised


def test_graph_replayed_minibatches_match_eager_minibatches():
    """PPOLearner replays a captured CUDA graph per minibatch (after two eager minibatches of that size); the weights after two
    updates must agree with a learner that runs every minibatch eagerly from the same initial weights on the same batches."""
    import copy
    from hhmarl_2d_b200 import VecLowLevelEnv, make_args, VecSampler, TorchPolicy, PPOLearner
    from hhmarl_2d_b200 import models as M
    torch.manual_seed(0)
    env = VecLowLevelEnv(256, make_args(level=3), device=0, seed=4)
    m1, m2 = M.build_policy_pair("fight")
    m1.cuda(); m2.cuda()
    e1, e2 = copy.deepcopy(m1), copy.deepcopy(m2)
    e2.shared_layer = e1.shared_layer                         # deepcopy of each model alone would split the shared layer
    smp = VecSampler(env, TorchPolicy(m1, 1), TorchPolicy(m2, 2), fragment_len=20, use_cuda_graph=False)
    lg = PPOLearner(m1, m2, num_sgd_iter=2, sgd_minibatch_size=1024, seed=3, use_cuda_graph=True)
    le = PPOLearner(e1, e2, num_sgd_iter=2, sgd_minibatch_size=1024, seed=3, use_cuda_graph=False)
    assert lg.use_graph and not le.use_graph
    assert torch.equal(lg.flat, le.flat)
    for _ in range(2):
        b = {k: v.clone() for k, v in smp.collect().items() if torch.is_tensor(v)}
        sg, se = lg.update(b), le.update(b)                   # the sampler's weights are NOT refreshed: same batches for both
        assert sg["minibatches"] == se["minibatches"] == 12   # 256 sequences, 51 per minibatch: 5 full + 1 remainder, twice
    assert 51 in lg._graphs                                   # the full-size minibatch was captured and replayed
    d = (lg.flat - le.flat).abs().max().item()
    moved = (le.flat - torch.cat([p.detach().reshape(-1) for p in le.params])).abs().max().item()
    assert moved == 0.0
    assert d < 2e-5, d
    for a, b_ in zip(sg["kl"] + sg["vf_loss"], se["kl"] + se["vf_loss"]):
        assert abs(a - b_) <= 1e-4 * max(1.0, abs(b_)), (sg, se)


def _nccl_worker(rank, world, port, out_dir):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from hhmarl_2d_b200 import VecLowLevelEnv, make_args, VecSampler, TorchPolicy, PPOLearner
    from hhmarl_2d_b200 import models as M
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    torch.manual_seed(0)                                      # identical initial weights
    m1, m2 = M.build_policy_pair("fight")
    m1.cuda(rank); m2.cuda(rank)
    n = 256
    env = VecLowLevelEnv(n, make_args(level=3), device=rank, seed=9, arena_base=rank * n)    # different arenas per rank
    smp = VecSampler(env, TorchPolicy(m1, 1), TorchPolicy(m2, 2), fragment_len=20, use_cuda_graph=False)
    learner = PPOLearner(m1, m2, num_sgd_iter=2, sgd_minibatch_size=1280)
    for _ in range(2):
        st = learner.update(smp.collect())
        smp.refresh_policy()
    torch.save({"flat": learner.flat.detach().cpu(), "rew": float(smp.buf["rew"].sum()), "mb": st["minibatches"]},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_nccl_ranks_end_an_update_with_identical_weights(tmp_path):
    """arenas sharded over ranks, ONE NCCL all-reduce of the flat gradient per minibatch.  Two ranks with
    different arenas (different rollouts) must hold bit-identical weights after the updates."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run by profiles/run_r2_multi.sh under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mp.spawn(_nccl_worker, args=(2, 29621, str(tmp_path)), nprocs=2, join=True)
    a, b = (torch.load(tmp_path / f"rank{r}.pt") for r in (0, 1))
    assert a["mb"] == b["mb"] == 8 and a["rew"] != b["rew"]      # different data ...
    assert torch.equal(a["flat"], b["flat"])                     # ... same weights


def test_ppo_update_on_gpu():
    from hhmarl_2d_b200 import VecLowLevelEnv, make_args, VecSampler, TorchPolicy, PPOLearner
    from hhmarl_2d_b200 import models as M
    torch.manual_seed(0)
    env = VecLowLevelEnv(256, make_args(level=3), device=0, seed=3)
    m1, m2 = M.build_policy_pair("fight")
    m1.cuda(); m2.cuda()
    smp = VecSampler(env, TorchPolicy(m1, 1), TorchPolicy(m2, 2), fragment_len=20, use_cuda_graph=False)
    learner = PPOLearner(m1, m2, num_sgd_iter=1, sgd_minibatch_size=1280)
    w0 = m1.act_out._model[0].weight.clone()
    st = learner.update(smp.collect())
    assert st["minibatches"] == 4 and np.isfinite(st["loss"]) and not torch.equal(w0, m1.act_out._model[0].weight)
    assert all(k >= 0 for k in st["kl"])


def test_native_multicategorical_sampler():
    """hh_sample_actions: frequencies follow softmax, logp matches torch, explore=0 is the per-head argmax,
    and the per-(arena, agent) counters make successive calls draw fresh, reproducible numbers."""
    from hhmarl_2d_b200 import _native as nat
    from hhmarl_2d_b200 import models as M
    from hhmarl_2d_b200.sampler import multicategorical_logp_entropy_kl
    torch.manual_seed(0)
    n = 40000
    row1, row2 = torch.randn(26), torch.randn(24)
    l1, l2 = row1.repeat(n, 1).cuda().contiguous(), row2.repeat(n, 1).cuda().contiguous()
    ctr = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
    act = torch.empty((n, 2, 4), dtype=torch.int32, device="cuda")
    logp = torch.empty((n, 2), device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def call(explore, counters):
        nat.check(nat.lib().hh_sample_actions(n, l1.data_ptr(), l2.data_ptr(), 123, 0, counters.data_ptr(), explore,
                                              act.data_ptr(), logp.data_ptr(), st), "hh_sample_actions")
        return act.clone(), logp.clone()

    a, lp = call(1, ctr)
    assert (ctr == 1).all() and (a[:, 1, 3] == 0).all()
    o = 0
    for h, w in enumerate((13, 9, 2, 2)):
        freq = torch.bincount(a[:, 0, h].long().cpu(), minlength=w).float() / n
        assert (freq - torch.softmax(row1[o:o + w], 0)).abs().max() < 0.012
        o += w
    ref1, _, _ = multicategorical_logp_entropy_kl(l1, a[:, 0, :], (13, 9, 2, 2))
    ref2, _, _ = multicategorical_logp_entropy_kl(l2, a[:, 1, :3], (13, 9, 2))
    assert torch.allclose(lp[:, 0], ref1, atol=2e-5) and torch.allclose(lp[:, 1], ref2, atol=2e-5)
    b, _ = call(1, ctr)
    assert not torch.equal(a, b)                                   # counters advanced -> new draws
    c, _ = call(1, torch.zeros_like(ctr))
    assert torch.equal(a, c)                                       # same counters -> same draws
    d, _ = call(0, ctr)
    assert (d[:, 0, :].long() == M.deterministic_actions(l1, 1)).all() and (d[:, 1, :3].long() == M.deterministic_actions(l2, 2)).all()


@pytest.mark.parametrize("mode", ["fight", "escape"])
def test_fused_policy_forward_matches_torch(mode):
    """csrc/hh_policy.cu (one launch: both policies' actor + central critic) against the fp32 torch forward of
    models.py and the packed cuBLAS forward: 3xTF32 reproduces fp32 to ~1e-5, plain TF32 to ~1e-2; ragged batch."""
    from hhmarl_2d_b200 import models as M
    from hhmarl_2d_b200.fused_forward import FusedPolicyPair, PackedPolicyPair
    torch.manual_seed(1)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        m1, m2 = M.build_policy_pair(mode)
        m1.cuda(); m2.cuda()
        for m in (m1, m2):                       # non-trivial biases (the initialiser sets them to zero)
            for prm in m.parameters():
                if prm.dim() == 1:
                    torch.nn.init.normal_(prm, std=0.1)
        B = 1000 + 37
        d = m1.central_dim
        f1 = torch.rand(B, d, device="cuda")
        f2 = torch.rand(B, d, device="cuda")
        f1[:, :7] = 0; f2[:, :7] = 0            # the sampler feeds zero actions ...
        f1[5:50, :7] = torch.rand(45, 7, device="cuda")   # ... the learner real ones
        with torch.no_grad():
            rl1, rv1 = m1.forward_flat(f1)
            rl2, rv2 = m2.forward_flat(f2)
        ref = (rl1, rv1, rl2, rv2)
        pk = PackedPolicyPair(m1, m2).forward(f1, f2)
        for a, b in zip(pk, ref):
            assert torch.allclose(a, b, atol=2e-5, rtol=1e-5)
        for prec, atol in ((2, 3e-5), (0, 3e-5), (1, 2e-2)):     # 2: tcgen05 / TMEM path (csrc/hh_policy_tc.cu)
            fu = FusedPolicyPair(m1, m2, precision=prec)
            out = [o.clone() for o in fu.forward(f1, f2)]
            for name, a, b in zip(("logits1", "value1", "logits2", "value2"), out, ref):
                assert a.shape == b.shape and torch.isfinite(a).all()
                err = (a - b).abs().max().item()
                assert err <= atol, (mode, prec, name, err)
            # weights changed by the learner -> refresh() re-packs in place
            with torch.no_grad():
                m1.act_out._model[0].weight.mul_(0.5)
                m1.shared_layer._model[0].bias.add_(0.01)
            fu.refresh()
            with torch.no_grad():
                nl1, nv1 = m1.forward_flat(f1)
            out = fu.forward(f1, f2)
            assert (out[0] - nl1).abs().max().item() <= atol and (out[1] - nv1).abs().max().item() <= atol
            with torch.no_grad():
                m1.act_out._model[0].weight.mul_(2.0)
                m1.shared_layer._model[0].bias.sub_(0.01)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("variant,mode_id", [("pair", 1), ("m64", 0)])
def test_tc_kernel_variants_match_torch(variant, mode_id):
    """HH_TC_MODE=pair: the CTA-pair form of the tcgen05 forward (cta_group::2 MMAs, each CTA streams half of the weight
    columns, relay + multicast commits); HH_TC_MODE=m64: 64-row tiles with hi and lo activations in shared memory.  Neither
    is the default (128-row tiles, lo halves in tensor memory; profiles/README.md, round 2); both stay tested variants.  Runs
    in a subprocess because the layout of the packed images is fixed when the library is loaded."""
    import os
    import subprocess
    import sys
    code = (
        "import torch\n"
        "from hhmarl_2d_b200 import _native as nat, models as M\n"
        "from hhmarl_2d_b200.fused_forward import FusedPolicyPair, FusedActor, run_chains\n"
        f"assert nat.lib().hh_policy_tc_mode() == {mode_id}\n"
        "torch.manual_seed(1)\n"
        "worst = 0.0\n"
        "for mode in ('fight', 'escape'):\n"
        "    m1, m2 = M.build_policy_pair(mode); m1.cuda(); m2.cuda()\n"
        "    for m in (m1, m2):\n"
        "        for p in m.parameters():\n"
        "            if p.dim() == 1: torch.nn.init.normal_(p, std=0.1)\n"
        "    for B in (1037, 64, 1):\n"
        "        f1 = torch.rand(B, m1.central_dim, device='cuda'); f2 = torch.rand(B, m1.central_dim, device='cuda')\n"
        "        with torch.no_grad():\n"
        "            ref = (*m1.forward_flat(f1), *m2.forward_flat(f2))\n"
        "        out = FusedPolicyPair(m1, m2, precision=2).forward(f1, f2)\n"
        "        worst = max(worst, max((a - b).abs().max().item() for a, b in zip(out, ref)))\n"
        "    fa = FusedActor(m1)\n"
        "    x = torch.rand(300, 30, device='cuda'); o = torch.empty(300, fa.n_out, device='cuda')\n"
        "    run_chains([lambda c: fa.fill_chain(c, x, 300, out=o)], torch.device('cuda'), 2)\n"
        "    with torch.no_grad():\n"
        "        worst = max(worst, (o - m1.actor(x[:, :fa.d_in])).abs().max().item())\n"
        "print('WORST', worst)\n"
        "assert worst < 3e-5\n")
    env = dict(os.environ, HH_TC_MODE=variant)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "WORST" in r.stdout


def test_policy_pack_image_layout():
    """hh_policy_pack: fp16 hi / lo halves of 2^s w in the K-major canonical layout, ring stage after ring stage; for the
    CTA-pair kernel every stage is split into the two CTAs' column halves (csrc/hh_policy_tc.cu)."""
    from hhmarl_2d_b200 import _native as nat
    from hhmarl_2d_b200.fused_forward import _pack_image
    torch.manual_seed(0)
    split = 1 + nat.lib().hh_policy_tc_pair()
    for (K, ld, n_cols, n_total, n_chunk, shift, ksteps, kps) in ((72, 512, 512, 512, 256, 0, 5, 1), (152, 152, 150, 160, 160, 6, 10, 1),
                                                                  (504, 32, 32, 32, 32, 0, 32, 8)):
        w = (torch.randn(K, ld, device="cuda") * 0.05).contiguous()
        w[3, 5] = 0.0
        img, us = _pack_image(w, n_cols, n_total, n_chunk, shift, ksteps, kps)
        torch.cuda.synchronize()
        unscale, scale = us.tolist()
        assert 8192 <= w[:, :n_cols].abs().max().item() * scale < 16384 and abs(unscale * scale * 4096 - 1) < 1e-12
        n_sub, n_stage = n_chunk // split, ksteps // kps
        # [chunk][stage][cta][kstep in stage][hi|lo][khalf][column of the CTA's half][kk]
        h = img.view(torch.float16).view(n_total // n_chunk, n_stage, split, kps, 2, 2, n_sub, 8).float()
        rec = (h[:, :, :, :, 0] + h[:, :, :, :, 1]) / scale                      # [chunk][stage][cta][j][khalf][cs][kk]
        rec = rec.permute(1, 3, 4, 6, 0, 2, 5).reshape(ksteps * 16, n_total)    # [(stage, j, khalf, kk)][(chunk, cta, cs)]
        want = torch.zeros(ksteps * 16, n_total, device="cuda")
        rows = min(K, ksteps * 16 - shift)
        want[shift:shift + rows, :n_cols] = w[:rows, :n_cols]
        assert (rec - want).abs().max().item() <= 2.0 ** -21 * w.abs().max().item()
        assert (h[:, :, :, :, 1].abs() <= h[:, :, :, :, 0].abs() * 2.0 ** -10 + 1e-30).all()     # lo is the rounding residue of hi


@pytest.mark.parametrize("precision", [2, 0])
def test_fused_actor_chains_gather_range_and_argmax(precision):
    """hh_policy_forward_ex: the four frozen-actor kinds (Fight1/2, Esc1/2 .actor) as chains of one launch, with a
    row gather, a device-side {begin, count} range and the per-head argmax epilogue, against the torch modules."""
    from hhmarl_2d_b200 import models as M
    from hhmarl_2d_b200.fused_forward import FusedActor, run_chains
    torch.manual_seed(3)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        f1, f2 = M.build_policy_pair("fight")
        e1, e2 = M.build_policy_pair("escape")
        models = [m.cuda().eval() for m in (f1, f2, e1, e2)]
        for m in models:
            for prm in m.parameters():
                if prm.dim() == 1:
                    torch.nn.init.normal_(prm, std=0.1)
        n = 3000
        x = torch.rand(n, 30, device="cuda")
        fas = [FusedActor(m) for m in models]
        # chain j works on its own quarter of a shuffled row list, given once as a host count and once as a device range
        perm = torch.randperm(n, device="cuda").to(torch.int32)
        cuts = [0, 700, 1500, 2300, n]
        outs = [torch.full((n, fa.n_out), float("nan"), device="cuda") for fa in fas]
        acts = torch.full((n, 4), -1, dtype=torch.int32, device="cuda")
        ranges = torch.tensor([[cuts[j], cuts[j + 1] - cuts[j]] for j in range(4)], dtype=torch.int32, device="cuda")
        fills = []
        for j, fa in enumerate(fas):
            if j % 2 == 0:
                rows = perm[cuts[j]:cuts[j + 1]].contiguous()
                fills.append(lambda c, fa=fa, rows=rows, j=j: fa.fill_chain(c, x, rows.numel(), out=outs[j], act_out=acts, rows=rows))
            else:
                fills.append(lambda c, fa=fa, j=j: fa.fill_chain(c, x, n, out=outs[j], act_out=acts, rows=perm, range_dev=ranges[j]))
        run_chains(fills, torch.device("cuda"), precision)
        n_checked = 0
        for j, (m, fa) in enumerate(zip(models, fas)):
            rows = perm[cuts[j]:cuts[j + 1]].long()
            with torch.no_grad():
                ref = m.actor(x[rows][:, :fa.d_in])
            got = outs[j][rows]
            assert (got - ref).abs().max().item() < 3e-5, j
            other = torch.ones(n, dtype=torch.bool, device="cuda"); other[rows] = False
            assert torch.isnan(outs[j][other]).all()                 # rows of other chains untouched
            want = M.deterministic_actions(ref, m.ac_type)
            top = torch.stack([torch.topk(ref[:, o:o + w], 2).values for o, w in
                               zip(np.cumsum((0,) + fa.splits[:-1]), fa.splits)], 1)     # [rows, heads, 2]
            clear = ((top[..., 0] - top[..., 1]) > 1e-3).all(1)
            a = acts[rows][:, :len(fa.splits)].long()
            assert torch.equal(a[clear], want[clear])
            if len(fa.splits) == 3:
                assert (acts[rows][:, 3] == 0).all()
            n_checked += int(clear.sum())
        assert n_checked > 2000
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("level,mode", [(4, "fight"), (5, "fight")])
def test_fused_opponents_match_torch_opponents(level, mode):
    from hhmarl_2d_b200.opponents import OpponentPolicies
    torch.manual_seed(4)
    n = 2500
    opp = OpponentPolicies(level, mode, seed=0, device="cuda")
    o3, o4 = torch.rand(n, 30, device="cuda"), torch.rand(n, 29, device="cuda")
    pset = torch.randint(3, 6, (n,), device="cuda").to(torch.uint8)
    o3[pset != 5, 26:] = 0; o4[pset != 5, 24:] = 0
    ref, logits = opp.act(o3, o4, pset, return_logits=True)
    got = opp.act_fused(o3, o4, pset)
    assert got.shape == ref.shape and got.dtype == torch.int32
    agree = (got == ref).all(2).all(1).float().mean().item()
    assert agree > 0.995, agree                                     # differences only at near-ties of a head


def test_sampler_graph_sees_refreshed_weights():
    """After a learner update `refresh_policy()` re-packs the fused kernel's weight buffers IN PLACE, so the captured
    CUDA graph samples with the new weights: recorded logits / values equal a torch forward of the updated modules."""
    from hhmarl_2d_b200 import VecLowLevelEnv, make_args, VecSampler, TorchPolicy
    from hhmarl_2d_b200 import models as M
    torch.manual_seed(5)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        env = VecLowLevelEnv(300, make_args(level=3), device=0, seed=13)
        m1, m2 = M.build_policy_pair("fight")
        m1.cuda(); m2.cuda()
        smp = VecSampler(env, TorchPolicy(m1, 1), TorchPolicy(m2, 2), fragment_len=6, use_cuda_graph=True)
        for _ in range(3):
            smp.collect()                                   # warm-up, capture, replay

        def check():
            b = smp.collect()
            for t in (0, 5):
                f1, f2 = b["flat1"][t].clone(), b["flat2"][t].clone()
                f1[:, :7] = 0; f2[:, :7] = 0               # sampling saw zero actions (the write-back came later)
                with torch.no_grad():
                    l1, v1 = m1.forward_flat(f1)
                    l2, v2 = m2.forward_flat(f2)
                assert (b["logits1"][t] - l1).abs().max().item() < 5e-5 and (b["logits2"][t] - l2).abs().max().item() < 5e-5
                assert (b["vf"][t][:, 0] - v1).abs().max().item() < 5e-5 and (b["vf"][t][:, 1] - v2).abs().max().item() < 5e-5

        check()
        with torch.no_grad():
            for m in (m1, m2):
                for prm in m.parameters():
                    prm.add_(0.02 * torch.randn_like(prm))
        smp.refresh_policy()
        check()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
