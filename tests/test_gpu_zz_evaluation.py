"""GPU tests of the evaluation helpers around the env kernels: evaluation statistics
of the commander environment (hhmarl_2d_b200.evaluation, csrc/hh_hier.cu eval_info_kernel) and the trajectory
recorder (hhmarl_2d_b200.trace).  The per-step info counters themselves are compared with the oracle in
tests/test_gpu_hier.py::test_hier_matches_oracle."""
import numpy as np
import pytest
import torch

from test_gpu_parity import _close, _vec

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kw", [{"glob_frac": 0.3, "hier_action_assess": False, "hier_opp_fight_ratio": 40},
                                {"friendly_kill": False, "rew_scale": 2.0}])
def test_eval_info_matches_oracle_other_configs(kw):
    """The info counters of step() under args.eval_info (env_base.py:91-107) against the oracle, every arena and step, for
    the non-default configurations of tests/test_gpu_hier.py::test_hier_matches_oracle (which checks the default one)."""
    from test_gpu_hier import _hier_vs_oracle
    _hier_vs_oracle(kw, check_eval=True, T=24)


def test_evaluate_commander(tmp_path):
    """evaluation.py on all arenas at once: the sums are consistent, the report has the reference's keys
    (postprocess_eval, evaluation.py:66-82), and the no-commander mode attacks the closest opponent only."""
    import json
    from hhmarl_2d_b200 import models as M
    from hhmarl_2d_b200.env_hier import VecHighLevelEnv, make_hier_args
    from hhmarl_2d_b200.evaluation import evaluate
    n = 256
    with pytest.raises(ValueError):
        evaluate(VecHighLevelEnv(8, make_hier_args(horizon=60), device=0, seed=1), None, 8)
    model = M.CommanderGru()
    M.fill_from_seed(model, 5, scale=1.0)   # a commander that uses all three actions
    model.cuda().eval()
    for m in (model, None):
        env = VecHighLevelEnv(n, make_hier_args(horizon=100, eval_info=True), device=0, seed=9)
        st = evaluate(env, m, n_episodes=500)
        ev = st.totals()
        assert ev["episodes"] == 2 * n                       # ceil(500 / 256) episodes per arena
        assert ev["agents_win"] + ev["opps_win"] + ev["draw"] <= ev["episodes"]
        assert ev["agent_fight"] + ev["agent_escape"] == ev["agent_steps"] and ev["opp_fight"] + ev["opp_escape"] == ev["opp_steps"]
        assert ev["opp1"] + ev["opp2"] + ev["opp3"] == ev["agent_fight"] and ev["opp3"] == 0
        assert 0 < ev["agent_steps"] <= 3 * ev["total_n_actions"] and 0 < ev["opp_steps"] <= 3 * ev["total_n_actions"]
        if m is None:
            assert ev["agent_escape"] == 0 and ev["opp1"] == ev["agent_fight"]
        path = tmp_path / "Metrics_Commander_3-vs-3.json"
        rep = st.save(str(path))
        assert list(json.loads(path.read_text())) == ["win", "lose", "draw", "fight", "esc", "fight_opp", "esc_opp", "opp1", "opp2", "opp3"]
        assert abs(rep["fight"] + rep["esc"] - 100) < 1e-9 and abs(rep["fight_opp"] + rep["esc_opp"] - 100) < 1e-9


def test_trace_recorder_on_device_env():
    """hhmarl_2d_b200.trace.TraceRecorder over the device API (torch `done`, hh_get_state read-back): samples equal
    the oracle's per-tick aircraft states (the recorder's rules are pinned on the CPU in tests/test_trace_cpu.py)."""
    import oracle as orc
    from hhmarl_2d_b200.trace import TraceRecorder
    n, T, level, seed, base = 64, 60, 2, 515, 40
    env = _vec(n, level, "fight", seed, arena_base=base, autoreset=True)
    watch = [0, 17, 63]
    oracles = {k: orc.OracleEnv(orc.make_args(level=level, agent_mode="fight"), seed, base + k) for k in watch}
    env.reset()
    rec = TraceRecorder(env, watch)
    rec.start()
    want = {k: {u: [] for u in range(4)} for k in watch}
    for k, o in oracles.items():
        o.reset()
        s = o.state()
        for u in range(4):
            want[k][u].append((s.steps, s.lat[u], s.lon[u], s.heading[u], s.speed[u]))
    rng = np.random.default_rng(6)
    open_ep = {k: True for k in watch}
    for t in range(T):
        act = np.stack([rng.integers(0, 13, (n, 2)), rng.integers(0, 9, (n, 2)), rng.integers(0, 2, (n, 2)),
                        rng.integers(0, 2, (n, 2))], axis=-1).astype(np.int32)
        *_, done = env.step(torch.from_numpy(act).cuda())
        rec.after_step(done)
        for k, o in oracles.items():
            if not open_ep[k]:
                continue
            *_, d = o.step(act[k])
            if d:
                open_ep[k] = False       # first episode only
                continue
            s = o.state()
            for u in range(4):
                if s.alive[u]:
                    want[k][u].append((s.steps, s.lat[u], s.lon[u], s.heading[u], s.speed[u]))
    for k in watch:
        ep = rec.episodes(k)[0]
        for u in range(4):
            w = np.asarray(want[k][u], np.float64).reshape(-1, 5)
            assert ep["units"][u + 1].shape == w.shape and len(w) >= 2
            _close(ep["units"][u + 1], w, f"arena {k} unit {u + 1}")


def test_hier_trace_recorder_on_device_env():
    """HierTraceRecorder on the real commander env: one sample per simulator tick and live aircraft (time stamps 0, 1, 2 ...
    without gaps, as many ticks as the env reports sub-steps), the last sample of an open episode is the aircraft's current
    state (the recorder's bookkeeping itself is tested on the CPU, tests/test_trace_cpu.py)."""
    from hhmarl_2d_b200.env_hier import VecHighLevelEnv, make_hier_args
    from hhmarl_2d_b200.trace import HierTraceRecorder
    n, watch = 32, [0, 5, 31]
    env = VecHighLevelEnv(n, make_hier_args(horizon=120), device=0, seed=21, autoreset=True)
    env.reset()
    rec = HierTraceRecorder(env, watch)
    rec.start()
    g = torch.Generator().manual_seed(3)
    ticks = {a: 0 for a in watch}
    n_done = 0
    for t in range(14):
        _, _, done = env.step(torch.randint(0, 3, (n, 3), generator=g).to(torch.int32).cuda())
        sub = env.substeps.cpu().numpy()
        d = done.cpu().numpy()
        rec.after_step(done)
        for a in watch:
            ticks[a] = 0 if d[a] else ticks[a] + int(sub[a])
            n_done += int(d[a])
    assert n_done >= 2
    st = env.get_state()
    for a in watch:
        eps = rec.episodes(a)
        assert len(eps) >= 1
        for i, ep in enumerate(eps):
            longest = max(len(v) for v in ep["units"].values())
            for u, v in ep["units"].items():
                assert len(v) >= 1 and v[0, 0] == 0 and (np.diff(v[:, 0]) == 1).all()
                assert ((v[:, 1] >= 4.9) & (v[:, 1] <= 5.6) & (v[:, 2] >= 6.9) & (v[:, 2] <= 7.6)).all()
            if i == len(eps) - 1:      # the open episode: as many ticks as the env counted, last sample = current state
                assert longest == ticks[a] + 1 == st[a].steps + 1
                for u, v in ep["units"].items():
                    if st[a].alive[u - 1]:
                        assert len(v) == longest and v[-1, 1] == st[a].lat[u - 1] and v[-1, 3] == st[a].hdg[u - 1]
