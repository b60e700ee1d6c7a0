"""hhmarl_2d_b200.trace.TraceRecorder (the reference's sim.trace_record_units, cmano_simulator.py:126-162) on the CPU:
the recorder reads the env through the host C ABI, here served by the emulation of the v4 step schedule
(tests/emu), and is compared with the oracle's per-tick states and -- where /root/reference is present -- with the
trace lists of the unmodified reference simulator."""
import json

import numpy as np
import pytest

import oracle as orc
import ref_harness as rh
import test_gpu_parity as tgp
from emu import emu_env
from test_emu_v4 import pursuit_actions

from hhmarl_2d_b200.trace import COLUMNS, TraceRecorder


def _expected_sample(s, u):
    return (float(s.steps), s.lat[u], s.lon[u], s.heading[u], s.speed[u])


def _run(autoreset, n=5, T=340, level=3, seed=4242, base=70):
    """Returns (recorder, expected): expected[a] = list of episodes, each {unit_id: [samples]} built from the
    oracle's state after every step by the reference's rule (a sample while the aircraft exists)."""
    with emu_env.emulated():
        env = tgp._vec(n, level, "fight", seed, arena_base=base, autoreset=autoreset)
        oracles = [orc.OracleEnv(orc.make_args(level=level, agent_mode="fight"), seed, base + k) for k in range(n)]
        env.reset_host()
        rec = TraceRecorder(env, range(n))
        rec.start()
        expected = [[] for _ in range(n)]
        cur = []
        for o in oracles:
            o.reset()
            s = o.state()
            cur.append({u + 1: [_expected_sample(s, u)] for u in range(4)})
        rng = np.random.default_rng(5)
        for t in range(T):
            act = pursuit_actions([o.state() for o in oracles], rng)
            _, _, _, gdone = env.step_host(act)
            mask = np.zeros(n, np.uint8)
            for k, o in enumerate(oracles):
                *_, d = o.step(act[k])
                assert bool(gdone[k]) == d
                s = o.state()
                if not (d and autoreset):        # with auto-reset the tick that ends the episode is not visible
                    for u in range(4):
                        if s.alive[u]:
                            cur[k][u + 1].append(_expected_sample(s, u))
                if d:
                    expected[k].append(cur[k])
                    o.reset()
                    s = o.state()
                    cur[k] = {u + 1: [_expected_sample(s, u)] for u in range(4)}
                    mask[k] = 1
            rec.after_step(gdone)
            if mask.any() and not autoreset:
                env.reset_host(mask)
                rec.after_reset(mask)
        for k in range(n):
            expected[k].append(cur[k])           # the open episode
        return rec, expected, env


@pytest.mark.parametrize("autoreset", [False, True])
def test_trace_matches_oracle_states(autoreset, tmp_path):
    rec, expected, env = _run(autoreset)
    n_short = 0
    for a in range(5):
        eps = rec.episodes(a)
        assert len(eps) == len(expected[a]) >= 2
        for i, (ep, want) in enumerate(zip(eps, expected[a])):
            assert ep["truncated_last_tick"] == (autoreset and i < len(eps) - 1)
            for u in range(1, 5):
                w = np.asarray(want[u], np.float64).reshape(-1, 5)
                g = ep["units"][u]
                assert g.shape == w.shape, (a, i, u)
                assert (g[:, 0] == w[:, 0]).all() and (np.diff(g[:, 0]) == 1).all() and g[0, 0] == 0
                tgp._close(g[:, 1:], w[:, 1:], f"arena {a} episode {i} unit {u}")
            lens = [len(ep["units"][u]) for u in range(1, 5)]
            n_short += min(lens) < max(lens)
    assert n_short >= 3          # aircraft that were shot down stop being traced before the episode ends
    out = rec.export_json(str(tmp_path / "trace.json"))
    back = json.loads((tmp_path / "trace.json").read_text())
    assert back["columns"] == list(COLUMNS) and back["map"]["right_lon"] == pytest.approx(7.3)
    assert sorted(back["arenas"]) == [str(a) for a in range(5)]
    assert back["arenas"]["2"][0]["units"]["3"] == out["arenas"]["2"][0]["units"]["3"] == rec.episodes(2)[0]["units"][3].tolist()
    with pytest.raises(ValueError):
        TraceRecorder(env, [5])


@pytest.mark.skipif(not rh.reference_available(), reason="/root/reference not present")
def test_trace_matches_reference_trace_lists():
    """One arena, first episode: the samples equal sim.trace_record_units of the unmodified reference simulator
    (positions, heading, speed; time as tick count).  An aircraft removed by the out-of-bounds check of _get_rewards
    (env_hetero.py:188-196) AFTER do_tick keeps that tick's sample in the reference and not here."""
    level, seed, base = 3, 99, 12
    with emu_env.emulated():
        env = tgp._vec(1, level, "fight", seed, arena_base=base, autoreset=False)
        ref = rh.ReferenceEnv(rh.make_namespace(level=level), seed, base)
        o = orc.OracleEnv(orc.make_args(level=level, agent_mode="fight"), seed, base)
        env.reset_host(); ref.reset(); o.reset()
        rec = TraceRecorder(env, [0])
        rec.start()
        rng = np.random.default_rng(8)
        for t in range(400):
            act = pursuit_actions([o.state()], rng)
            _, _, _, gdone = env.step_host(act)
            *_, d = ref.step(act[0])
            o.step(act[0])
            rec.after_step(gdone)
            assert bool(gdone[0]) == d
            if d:
                break
        assert d
        ep = rec.episodes(0)[0]
        tr = ref.env.sim.trace_record_units
        t0 = tr[1][0][0]
        n_cmp = 0
        for u in range(1, 5):
            want = np.array([[(tm - t0).total_seconds(), p.lat, p.lon, h, s] for tm, p, h, s in tr[u]], np.float64)
            got = ep["units"][u]
            assert len(want) - 1 <= len(got) <= len(want), (u, len(got), len(want))
            tgp._close(got, want[:len(got)], f"unit {u}")
            n_cmp += len(got)
        assert n_cmp > 100
