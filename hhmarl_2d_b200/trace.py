"""Trajectory recording / export for a separate (CPU) renderer.

The reference keeps, per aircraft, `sim.trace_record_units[unit_id]` = [(utc_time, position, heading, speed), ...]:
one entry when the aircraft is registered (env_base.py:581 -> cmano_simulator.py:126-131) and one after every
tick while the aircraft exists (cmano_simulator.py:148-150, 159-162); rockets are not traced.  Its plotter draws
these poly-lines (env_base.py:587-607).  Plotting is out of scope here; this module records the same samples for
selected arenas of a VecLowLevelEnv from the env's state (hh_get_state, include/hhmarl_b200.h) and writes them to a
file.  It reads the whole state back every tick, so it is an evaluation tool, not part of the rollout path.

One difference: an aircraft removed by the out-of-bounds check of _get_rewards (env_hetero.py:188-196), which runs
AFTER do_tick, keeps that tick's sample in the reference; here the state is read after the whole step, so that last
sample is absent (tests/test_trace_cpu.py compares with the reference's own lists up to it).

Time stamps are seconds since the start of the episode (tick = 1 s, cmano_simulator.py:80); the reference stamps
wall-clock `datetime.now()` + ticks, of which only the differences mean anything.
"""
from __future__ import annotations

import json

import numpy as np

N_AIRCRAFT = 4   # 2-vs-2: ids 1, 2 agents, 3, 4 opponents (env_base.py:556-560)
COLUMNS = ("t", "lat", "lon", "heading", "speed")


class TraceRecorder:
    """rec = TraceRecorder(env, arenas); env.reset(); rec.start(); loop: env.step(...); rec.after_step(done).

    With `autoreset=False` envs call `rec.after_reset(mask)` after resetting the finished arenas; the traces are
    then complete up to and including the tick that ended the episode.  With auto-reset the state read after that
    tick already belongs to the next episode: the finished episode's trace ends one tick early (flag
    `truncated_last_tick` in the export) and the new one starts with its initial sample."""

    def __init__(self, env, arenas, autoreset=None):
        self.env = env
        self.arenas = [int(a) for a in arenas]
        if any(a < 0 or a >= env.n_arenas for a in self.arenas):
            raise ValueError("TraceRecorder: arena index out of range")
        self.autoreset = bool(env._cfg.autoreset) if autoreset is None else bool(autoreset)
        self.finished = {a: [] for a in self.arenas}    # arena -> list of episodes
        self._cur = {a: None for a in self.arenas}

    @staticmethod
    def _new_episode():
        return {"units": {u + 1: [] for u in range(N_AIRCRAFT)}, "truncated_last_tick": False}

    def _sample(self, st, a):
        ep = self._cur[a]
        t = float(st["steps"][a])
        for u in range(N_AIRCRAFT):
            if st["alive"][a, u]:   # _store_unit_state: only while unit_exists (cmano_simulator.py:159-162)
                ep["units"][u + 1].append((t, float(st["lat"][a, u]), float(st["lon"][a, u]), float(st["heading"][a, u]),
                                           float(st["speed"][a, u])))

    def _close(self, a, truncated):
        ep = self._cur[a]
        if ep is not None:
            ep["truncated_last_tick"] = truncated
            self.finished[a].append(ep)
        self._cur[a] = None

    def start(self):
        """After env.reset(): every selected arena begins an episode with its registration sample."""
        st = self.env.get_state()
        for a in self.arenas:
            self._close(a, False)
            self._cur[a] = self._new_episode()
            self._sample(st, a)

    def after_step(self, done):
        """After env.step(): `done` is the step's done vector (numpy / torch / list, length n_arenas)."""
        done = np.asarray(done.cpu() if hasattr(done, "cpu") else done).astype(bool)
        st = self.env.get_state()
        for a in self.arenas:
            if self._cur[a] is None:
                continue                                  # finished, waiting for after_reset
            if done[a] and self.autoreset:
                self._close(a, True)
                self._cur[a] = self._new_episode()
                self._sample(st, a)
            else:
                self._sample(st, a)
                if done[a]:
                    self._close(a, False)

    def after_reset(self, mask=None):
        """After a (masked) env.reset() of an autoreset=False env: the reset arenas begin new episodes."""
        mask = None if mask is None else np.asarray(mask.cpu() if hasattr(mask, "cpu") else mask).astype(bool)
        st = self.env.get_state()
        for a in self.arenas:
            if mask is None or mask[a]:
                self._close(a, False)
                self._cur[a] = self._new_episode()
                self._sample(st, a)

    def episodes(self, arena, include_open=True):
        """List of episodes of one arena: {"units": {unit_id: float64 [k, 5] (COLUMNS)}, "truncated_last_tick": bool}."""
        eps = list(self.finished[int(arena)])
        if include_open and self._cur[int(arena)] is not None:
            eps.append(self._cur[int(arena)])
        return [{"units": {u: np.asarray(rows, np.float64).reshape(-1, len(COLUMNS)) for u, rows in ep["units"].items()},
                 "truncated_last_tick": ep["truncated_last_tick"]} for ep in eps]

    def export_json(self, path, include_open=True):
        """{"columns": [...], "map": {...}, "arenas": {arena: [episode, ...]}} -- plain lists, one file per run."""
        a = self.env.args
        out = {"columns": list(COLUMNS),
               # MapLimits(7.0, 5.0, 7.0 + map_size, 5.0 + map_size) (env_base.py:43)
               "map": {"left_lon": 7.0, "bottom_lat": 5.0, "right_lon": 7.0 + float(a.map_size),
                       "top_lat": 5.0 + float(a.map_size)},
               "arenas": {str(ar): [{"units": {str(u): v.tolist() for u, v in ep["units"].items()},
                                     "truncated_last_tick": ep["truncated_last_tick"]}
                                    for ep in self.episodes(ar, include_open)] for ar in self.arenas}}
        with open(path, "w") as f:
            json.dump(out, f)
        return out


class HierTraceRecorder:
    """The same samples for selected arenas of a VecHighLevelEnv (3-vs-3, aircraft ids 1-6): the reference's evaluation
    (evaluation.py:62-63) plots the commander episodes from sim.trace_record_units, one sample per simulator tick.  A
    commander step makes up to 16 ticks inside `env.step`; the recorder hooks `env.tick_hook`, reads the arena records
    after every tick (hh_hier_get_state) and appends a sample for the arenas whose tick counter moved.

    rec = HierTraceRecorder(env, arenas); env.reset(); rec.start(); loop: env.step(actions); rec.after_step(env.done).
    With auto-reset the record read after the last tick of an episode is still the finished episode's (the reset
    happens in hh_hier_end), so the traces are complete; `after_step` then opens the next episode."""

    N_UNITS = 6

    def __init__(self, env, arenas):
        self.env = env
        self.arenas = [int(a) for a in arenas]
        if any(a < 0 or a >= env.n_arenas for a in self.arenas):
            raise ValueError("HierTraceRecorder: arena index out of range")
        self.finished = {a: [] for a in self.arenas}
        self._cur = {a: None for a in self.arenas}
        self._last_steps = {a: -1 for a in self.arenas}
        env.tick_hook = self._on_tick

    def _sample(self, st, a):
        rec = st[a]
        ep = self._cur[a]
        t = float(rec.steps)
        for u in range(self.N_UNITS):
            if rec.alive[u]:
                ep[u + 1].append((t, float(rec.lat[u]), float(rec.lon[u]), float(rec.hdg[u]), float(rec.spd[u])))
        self._last_steps[a] = int(rec.steps)

    def _open(self, st, a):
        self._cur[a] = {u + 1: [] for u in range(self.N_UNITS)}
        self._sample(st, a)

    def start(self):
        st = self.env.get_state()
        for a in self.arenas:
            if self._cur[a] is not None:
                self.finished[a].append(self._cur[a])
            self._open(st, a)

    def _on_tick(self, sub_step):
        st = self.env.get_state()
        for a in self.arenas:
            if self._cur[a] is not None and int(st[a].steps) != self._last_steps[a]:   # idle arenas do not tick
                self._sample(st, a)

    def after_step(self, done):
        done = np.asarray(done.cpu() if hasattr(done, "cpu") else done).astype(bool)
        if not done[self.arenas].any():
            return
        st = self.env.get_state()
        for a in self.arenas:
            if done[a] and self._cur[a] is not None:
                self.finished[a].append(self._cur[a])
                self._cur[a] = None
                if int(st[a].steps) == 0:          # auto-reset env: the next episode has begun
                    self._open(st, a)

    def after_reset(self, mask=None):
        mask = None if mask is None else np.asarray(mask.cpu() if hasattr(mask, "cpu") else mask).astype(bool)
        st = self.env.get_state()
        for a in self.arenas:
            if mask is None or mask[a]:
                if self._cur[a] is not None:
                    self.finished[a].append(self._cur[a])
                self._open(st, a)

    def episodes(self, arena, include_open=True):
        eps = list(self.finished[int(arena)])
        if include_open and self._cur[int(arena)] is not None:
            eps.append(self._cur[int(arena)])
        return [{"units": {u: np.asarray(rows, np.float64).reshape(-1, len(COLUMNS)) for u, rows in ep.items()},
                 "truncated_last_tick": False} for ep in eps]

    def export_json(self, path, include_open=True):
        ms = float(self.env.args.map_size)
        out = {"columns": list(COLUMNS),
               "map": {"left_lon": 7.0, "bottom_lat": 5.0, "right_lon": 7.0 + ms, "top_lat": 5.0 + ms},
               "arenas": {str(ar): [{"units": {str(u): v.tolist() for u, v in ep["units"].items()},
                                     "truncated_last_tick": False} for ep in self.episodes(ar, include_open)]
                          for ar in self.arenas}}
        with open(path, "w") as f:
            json.dump(out, f)
        return out
