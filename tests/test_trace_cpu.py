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


class _Rec:
    def __init__(self):
        self.steps, self.alive = 0, [1] * 6
        self.lat, self.lon, self.hdg, self.spd = [0.0] * 6, [0.0] * 6, [0.0] * 6, [0.0] * 6


class _FakeHierEnv:
    """Stands in for VecHighLevelEnv: arenas tick a varying number of times per commander step, idle afterwards, lose an
    aircraft now and then, and auto-reset in the end phase of the step that finishes the episode."""

    def __init__(self, n, horizon, autoreset, seed=0):
        from argparse import Namespace
        self.n_arenas, self.horizon, self.autoreset = n, horizon, autoreset
        self.args = Namespace(map_size=0.5)
        self.rng = np.random.default_rng(seed)
        self.tick_hook = None
        self.recs = [_Rec() for _ in range(n)]
        self.truth = [[] for _ in range(n)]     # per arena: list of episodes {unit: [samples]}
        self.done = np.zeros(n, np.uint8)

    def _log(self, a):
        r = self.recs[a]
        for u in range(6):
            if r.alive[u]:
                self.truth[a][-1][u + 1].append((float(r.steps), r.lat[u], r.lon[u], r.hdg[u], r.spd[u]))

    def _reset_arena(self, a):
        r = self.recs[a] = _Rec()
        r.lat = list(5.0 + self.rng.random(6) * 0.5); r.lon = list(7.0 + self.rng.random(6) * 0.5)
        r.hdg = list(self.rng.random(6) * 360); r.spd = list(100 + self.rng.random(6) * 500)
        self.truth[a].append({u + 1: [] for u in range(6)})
        self._log(a)

    def reset(self, mask=None):
        for a in range(self.n_arenas):
            if mask is None or mask[a]:
                self._reset_arena(a)

    def get_state(self):
        return self.recs

    def step(self, _actions=None):
        n_ticks = self.rng.integers(11, 17, self.n_arenas)
        finished = self.done.astype(bool) & (not self.autoreset)
        for s in range(16):
            for a in range(self.n_arenas):
                r = self.recs[a]
                if s < n_ticks[a] and not finished[a] and r.steps < self.horizon:
                    r.steps += 1
                    for u in range(6):
                        r.lat[u] += 1e-4; r.lon[u] -= 2e-4; r.hdg[u] = (r.hdg[u] + 3.0) % 360
                    if self.rng.random() < 0.03:
                        r.alive[int(self.rng.integers(0, 6))] = 0
                    self._log(a)
            if self.tick_hook is not None:
                self.tick_hook(s)
        for a in range(self.n_arenas):
            if finished[a]:
                continue
            self.done[a] = self.recs[a].steps >= self.horizon
            if self.done[a] and self.autoreset:
                self._reset_arena(a)
        return self.done


@pytest.mark.parametrize("autoreset", [True, False])
def test_hier_trace_recorder_follows_every_tick(autoreset, tmp_path):
    from hhmarl_2d_b200.trace import HierTraceRecorder
    env = _FakeHierEnv(5, horizon=70, autoreset=autoreset)
    env.reset()
    rec = HierTraceRecorder(env, [0, 2, 4])
    rec.start()
    for t in range(30):
        done = env.step()
        rec.after_step(done)
        if not autoreset and done.any():
            mask = done.copy()
            env.done[:] = 0
            env.reset(mask)
            rec.after_reset(mask)
    for a in (0, 2, 4):
        eps = rec.episodes(a)
        assert len(eps) == len(env.truth[a]) >= 4
        for ep, want in zip(eps, env.truth[a]):
            for u in range(1, 7):
                w = np.asarray(want[u], np.float64).reshape(-1, 5)
                assert ep["units"][u].shape == w.shape and np.array_equal(ep["units"][u], w), (a, u)
    out = rec.export_json(str(tmp_path / "hier_trace.json"))
    assert out["map"]["top_lat"] == 5.5 and sorted(out["arenas"]) == ["0", "2", "4"]
    assert json.loads((tmp_path / "hier_trace.json").read_text())["arenas"]["2"][1]["units"]["6"] == rec.episodes(2)[1]["units"][6].tolist()
