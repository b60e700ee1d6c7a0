"""GPU HighLevelEnv (csrc/hh_hier.cu via hh_hier_*) against the C oracle of envs/env_hier.py.  The oracle's
policy callback replays the actions the GPU path's batched networks chose (and checks the observation each
query was made with), so the environment is compared exactly and the networks separately."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6


def _close(a, b, what):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    err = np.abs(a - b) - (ATOL + RTOL * np.abs(b))
    assert (err <= 0).all(), f"{what}: max excess {err.max():.3e}"


@pytest.mark.parametrize("kw", [{}, {"glob_frac": 0.3, "hier_action_assess": False, "hier_opp_fight_ratio": 40},
                                {"friendly_kill": False, "rew_scale": 2.0}])
def test_hier_matches_oracle(kw):
    # the eval-info counters are compared here for the default configuration and, for the other two, in
    # tests/test_gpu_zz_evaluation.py (same harness, run last)
    _hier_vs_oracle(kw, check_eval=not kw)


def _hier_vs_oracle(kw, check_eval, T=40):
    import oracle as orc
    from hhmarl_2d_b200.env_hier import VecHighLevelEnv, make_hier_args
    n, seed, base = 48, 606, 300
    env = VecHighLevelEnv(n, make_hier_args(**kw), device=0, seed=seed, arena_base=base, autoreset=True)
    env.eval_info = True               # args.eval_info (env_base.py:91-107) only adds the info counters
    from hhmarl_2d_b200.evaluation import EvalStats
    stats, want = EvalStats(env.dev), np.zeros(12, np.int64)
    queue = [[] for _ in range(n)]     # per arena: FIFO of (unit_id, mode, ac_type, obs, action) the GPU produced

    def make_fn(k):
        def fn(unit_id, ac_type, mode, pset, obs):
            assert queue[k], f"arena {k}: oracle asks for a policy query the GPU path did not make"
            u, m, ty, gobs, act = queue[k].pop(0)
            assert (u, m, ty) == (unit_id, mode, ac_type), (k, (u, m, ty), (unit_id, mode, ac_type))
            _close(gobs[:len(obs)], obs, f"arena {k} unit {unit_id} low-level obs")
            assert (gobs[len(obs):] == 0).all()
            return act[:4 if ac_type == 1 else 3]
        return fn

    oracles = [orc.OracleHierEnv(orc.make_hier_args(**kw), seed, base + k, make_fn(k)) for k in range(n)]
    obs = env.reset().cpu().numpy()
    _close(obs, np.stack([o.reset() for o in oracles]), "reset obs")
    rng = np.random.default_rng(1)
    n_done = 0
    for t in range(T):
        ca = rng.integers(0, 3, (n, 3)).astype(np.int32)
        env.trace = []
        gobs, grew, gdone = env.step(torch.from_numpy(ca).cuda())
        gobs, grew, gdone, gsub = gobs.cpu().numpy(), grew.cpu().numpy(), gdone.cpu().numpy(), env.substeps.cpu().numpy()
        ginfo = env.info.cpu().numpy()
        stats.update(env.info, env.done)
        # rebuild, per arena, the sequence of policy queries in the reference's order: per sub-step units 1..6
        for k in range(n):
            queue[k].clear()
        for i in range(0, len(env.trace), 2):
            (_, oa, ia, _), (_, oo, io, act) = env.trace[i], env.trace[i + 1]
            for k in range(n):
                for u in range(6):
                    info, o = (ia, oa) if u < 3 else (io, oo)
                    if info[k, u] & 1:
                        queue[k].append((u + 1, (info[k, u] >> 1) & 1, 2 if info[k, u] & 4 else 1, o[k, u], act[k, u]))
        for k, o in enumerate(oracles):
            eo, er, ed, info = o.step(ca[k])
            assert not queue[k], f"arena {k}: GPU made {len(queue[k])} more policy queries than the oracle"
            assert bool(gdone[k]) == ed and gsub[k] == info[0], (t, k, gsub[k], info[0])
            _close(grew[k], er, f"t={t} arena={k} rew")
            if check_eval:
                ei = list(o.eval_info().values())   # before the reset, like the reference's step()
                assert list(ginfo[k]) == ei, (t, k, list(ginfo[k]), ei)
                want += np.asarray(ei)
            if ed:
                eo = o.reset()
                n_done += 1
            _close(gobs[k], eo, f"t={t} arena={k} obs")
        if t % 10 == 9:
            st = env.get_state()
            for k, o in enumerate(oracles):
                s = o.state()
                assert [st[k].steps, st[k].alive_ag, st[k].alive_op, st[k].next_id, st[k].dg, st[k].dc] == \
                       [s.steps, s.alive_agents, s.alive_opps, s.next_unit_id, s.draws_g, s.draws_c], (t, k)
                assert list(st[k].alive) == list(s.alive[:6]) and list(st[k].mrem) == list(s.missile_remain[:6])
                _close(list(st[k].lat), list(s.lat[:6]), "lat"); _close(list(st[k].hdg), list(s.heading[:6]), "hdg")
                assert st[k].err == 0
    if T >= 40:
        assert n_done >= n // 2
    if check_eval:
        tot = stats.totals()
        assert [tot[k] for k in orc.EVAL_INFO_KEYS] == list(want) and tot["episodes"] == n_done and tot["total_n_actions"] == n * T
        assert tot["agents_win"] + tot["opps_win"] + tot["draw"] <= n_done


def test_commander_sampler_fragment():
    from hhmarl_2d_b200.env_hier import CommanderSampler, VecHighLevelEnv
    from hhmarl_2d_b200 import models as M
    torch.manual_seed(0)
    env = VecHighLevelEnv(256, device=0, seed=5)
    model = M.CommanderGru().cuda()
    smp = CommanderSampler(env, model, fragment_len=6)
    b = smp.collect()
    n = 256
    assert b["flat"].shape == (6, n, 3, 105) and (b["actions"] >= 0).all() and (b["actions"] <= 2).all()
    # central observation: own obs then the two team-mates in id order (train_hier.py:134-165)
    f = b["flat"]
    assert torch.equal(f[:, :, 1, 3:37], f[:, :, 0, 37:71]) and torch.equal(f[:, :, 0, 3:37], f[:, :, 1, 37:71])
    assert torch.equal(f[:, :, 2, 37:71], f[:, :, 0, 3:37]) and torch.equal(f[:, :, 2, 71:105], f[:, :, 1, 3:37])
    # action write-back, own first then team-mates, scaled by 1 / N_OPP_HL
    a = b["actions"].float() / 2
    assert torch.equal(f[:, :, 0, 0], a[:, :, 0]) and torch.equal(f[:, :, 0, 1], a[:, :, 1]) and torch.equal(f[:, :, 1, 1], a[:, :, 0])
    assert torch.isfinite(b["adv"]).all() and (b["substeps"] >= 1).all() and (b["substeps"] <= 16).all()
    lsm = torch.log_softmax(b["logits"], -1).gather(3, b["actions"].long()[..., None])[..., 0]
    assert torch.allclose(lsm, b["logp"], atol=1e-5)


def test_graph_replayed_commander_steps_match_eager_steps():
    """VecHighLevelEnv replays one captured CUDA graph per commander step (after two eager steps): observations, rewards, done
    flags, sub-step counts and the eval-info counters must equal an eager env's, step by step, with episodes ending in between."""
    from hhmarl_2d_b200.env_hier import VecHighLevelEnv, make_hier_args
    n, T = 512, 12
    torch.manual_seed(3)
    ca = torch.randint(0, 3, (T, n, 3)).to(torch.int32).cuda()
    envs = []
    # one handle, eager | two groups of arenas on two streams inside the graph | three ragged groups, eager (phase by phase)
    for use_graph, groups in ((False, 1), (True, 2), (False, 3), (True, 1)):
        e = VecHighLevelEnv(n, make_hier_args(horizon=60, eval_info=True), device=0, seed=8, arena_base=70, autoreset=True,
                            groups=groups)
        assert e.groups == groups
        e.use_cuda_graph = use_graph
        envs.append(e)
    first = envs[0].reset().clone()
    for e in envs[1:]:
        assert torch.equal(first, e.reset())
    n_done = 0
    for t in range(T):
        outs = []
        for e in envs:
            o, r, d = e.step(ca[t])
            outs.append((o.clone(), r.clone(), d.clone(), e.substeps.clone(), e.info.clone()))
        for other in outs[1:]:
            for a, b in zip(outs[0], other):
                assert torch.equal(a, b), t
        n_done += int(outs[0][2].sum())
    assert envs[1]._graph is not None and envs[3]._graph is not None and envs[0]._graph is None and envs[2]._graph is None
    assert n_done > 0
    sa, sb = envs[0].get_state(), envs[1].get_state()
    assert bytes(sa) == bytes(sb)


def test_hier_full_size_properties():
    """BASELINE config 5 size (8 192 arenas, 3-vs-3): two runs agree bit for bit, observations stay in Box(0, 1),
    every arena makes between 1 and 16 sub-steps per commander step, a ragged slice of the arenas reproduces."""
    from hhmarl_2d_b200.env_hier import VecHighLevelEnv
    n, T = 8192, 3
    torch.manual_seed(2)
    ca = torch.randint(0, 3, (T, n, 3)).to(torch.int32).cuda()

    def run(n_arenas, base, sl):
        env = VecHighLevelEnv(n_arenas, device=0, seed=4, arena_base=base, autoreset=True)
        outs = [env.reset().clone().reshape(n_arenas, -1)]
        for t in range(T):
            o, r, d = env.step(ca[t, sl].contiguous())
            outs.append(torch.cat([o.reshape(n_arenas, -1), r, d.float()[:, None], env.substeps.float()[:, None]], dim=1).clone())
        return outs

    full = run(n, 0, slice(0, n))
    again = run(n, 0, slice(0, n))
    for a, b in zip(full, again):
        assert torch.equal(a, b)
    for o in full[1:]:
        assert torch.isfinite(o).all() and (o[:, :102] >= 0).all() and (o[:, :102] <= 1).all()
        assert (o[:, -1] >= 1).all() and (o[:, -1] <= 16).all()
    part = run(333, 5000, slice(5000, 5333))
    for a, b in zip(full, part):
        assert torch.equal(a[5000:5333], b)
