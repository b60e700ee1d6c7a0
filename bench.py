#!/usr/bin/env python
"""bench.py -- env-steps/s of the B200-native hhmarl_2D hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference path on the host cores

Workload (BASELINE.json configs[1]): 8 192 arenas of the 2-vs-2 level-3 (scripted-opponent) fight
scenario per GPU, all arenas advanced one tick per "step" by ONE fused kernel launch (action
decode -> scripted opponents -> tick -> rewards -> auto-reset -> observations).  Actions are
uniform MultiDiscrete samples, pre-generated on the device (`value`) or supplied from host
memory each step (`e2e`, through the C ABI's host entry point hh_step_host: H2D of actions and
D2H of observations/rewards/done inside the timed region).

  value     whole-job env-steps/s, inputs resident in HBM, per-step CUDA events, L2 flushed
            between steps, max over ranks
  e2e       same metric through hh_step_host (host buffers)
  roofline  algorithmic bytes of one launch / its CUDA-event duration vs the measured HBM peak
  cpu_baseline  the C oracle (port of the reference algorithm) on this box's host cores,
            bounded sample
The reference itself is Python + ray and cannot travel to the GPU box (no ray, no
geographiclib, no /root/reference there); `--impl reference` therefore times the oracle port
(oracle/hhmarl_oracle.c) with every host thread it can use, as the task's tier rules prescribe.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "env-steps/sec"
UNIT = "env-steps/s"
ALGO_BYTES_PER_ENV_STEP = 840          # SURVEY.md section 8(d), levels 1-3
FALLBACK_HBM_GBS = 6650.0              # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arenas", type=int, default=8192, help="arenas per GPU")
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--cpu-sample-steps", type=int, default=400_000, help="env-steps per host thread")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rollout", action="store_true")
    ap.add_argument("--no-hier", action="store_true")
    ap.add_argument("--no-l5", action="store_true")
    ap.add_argument("--no-ppo", action="store_true")
    ap.add_argument("--leg", default="step", choices=["step", "rollout"],
                    help="what `value` measures: 'step' = the fused env-step kernel with pre-generated actions (default; the "
                         "kernel the roofline is about), 'rollout' = the whole on-device sampler loop (policy forward + "
                         "sampling + env step + GAE); both arms accept it, and the default line carries both")
    ap.add_argument("--l5-arenas", type=int, default=0, help="arenas per GPU of the level-5 leg (0: 32 768 on one GPU = "
                    "BASELINE config 3, 8 192 per GPU under torchrun = config 4)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------ CPU legs
def cpu_oracle_throughput(level: int, steps_per_thread: int, threads: int):
    """env-steps/s of the C oracle, one arena per host thread (the reference's one-env-per-rollout-
    worker layout, train_hetero.py:212), uniformly random actions, auto-reset."""
    from concurrent.futures import ThreadPoolExecutor
    import oracle as orc
    envs = [orc.OracleEnv(orc.make_args(level=level), 0, k) for k in range(threads)]
    for e in envs:
        e.run_random(1000, 7)  # warm caches / page in
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:   # ctypes releases the GIL inside the C call
        list(ex.map(lambda e: e.run_random(steps_per_thread, 1), envs))
    dt = time.perf_counter() - t0
    return threads * steps_per_thread / dt, dt


def cpu_sampler_equivalent(level: int, seconds: float = 4.0):
    """SURVEY.md 8(d) CPU side-by-side (ii): what ONE rollout worker of the reference does per env step -- the env
    step (C oracle here, so this is an upper bound for the Python env) plus a batch-1 torch-CPU forward of both
    policies (actor + central critic, train_hetero.py:162-181 observation layout with zero team-mate actions) and
    MultiCategorical sampling -- on one host core.  Returns env-steps/s of that single worker."""
    import numpy as np
    import torch
    import oracle as orc
    from hhmarl_2d_b200 import models as M
    prev = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        torch.manual_seed(0)
        m1, m2 = M.build_policy_pair("fight")
        m1.eval(); m2.eval()
        env = orc.OracleEnv(orc.make_args(level=level), 0, 0)
        o1, o2 = env.reset()
        z1, z = torch.zeros(1, 4), torch.zeros(1, 3)

        def act(model, own, other, a_own, a_other, splits):
            d = {"obs_1_own": torch.from_numpy(own[None]), "obs_2": torch.from_numpy(other[None]), "act_1_own": a_own, "act_2": a_other}
            logits, _ = model({"obs": d}, [], None)
            model.value_function()
            out, o = [], 0
            for n in splits:
                out.append(int(torch.distributions.Categorical(logits=logits[0, o:o + n]).sample()))
                o += n
            return out

        n_steps, t0 = 0, time.perf_counter()
        with torch.no_grad():
            while time.perf_counter() - t0 < seconds:
                a1 = act(m1, o1, o2, z1, z, (13, 9, 2, 2))
                a2 = act(m2, o2, o1, z, z1, (13, 9, 2)) + [0]
                o1, o2, _, _, done = env.step(np.array([a1, a2], np.int32))
                if done:
                    o1, o2 = env.reset()
                n_steps += 1
        return n_steps / (time.perf_counter() - t0)
    finally:
        torch.set_num_threads(prev)


def _sampler_worker(args):
    level, seconds = args
    return cpu_sampler_equivalent(level, seconds)


def cpu_sampler_all_cores(level: int, seconds: float, workers: int):
    """The reference's sampler layout (train_hetero.py:212: one rollout worker per core): `workers` processes, each one
    reference-style worker (cpu_sampler_equivalent) for `seconds`.  Returns aggregate env-steps/s."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        pool.map(_sampler_worker, [(level, 0.2)] * workers)          # imports + warm-up outside the timed sample
        rates = pool.map(_sampler_worker, [(level, seconds)] * workers)
    return float(sum(rates)), rates


def best_thread_count(level: int):
    """Host boxes may expose more logical CPUs than the cgroup lets us use: probe a few thread
    counts on a small sample and keep the fastest."""
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cands = sorted({c for c in (4, 8, 16, 32, 64, 96, 128, 192, avail) if c <= avail} | {avail})
    best, best_v = cands[0], 0.0
    for c in cands:
        v, _ = cpu_oracle_throughput(level, 15_000, c)
        if v > best_v * 1.03:
            best, best_v = c, v
    best_thread_count.rate = best_v
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = best_thread_count(args.level)
    # bounded sample per bench step: the whole --steps/--warmup run stays within ~2 minutes on this box
    budget = 100.0 * best_thread_count.rate / max(1, args.steps + args.warmup) / cores
    per_step = int(min(max(20_000, args.cpu_sample_steps // 8), max(2_000, budget)))
    for _ in range(args.warmup):
        cpu_oracle_throughput(args.level, per_step // 10, cores)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        v, dt = cpu_oracle_throughput(args.level, per_step, cores)
        t_total += dt
        n_total += cores * per_step
    value = n_total / t_total
    sample = (f"{cores} host threads (best of a thread-count probe; box reports {os.cpu_count()} logical CPUs) x "
              f"{per_step} env-steps of the L{args.level} scenario per bench step, random actions")
    # the sampler leg of the reference arm: one reference-style rollout worker per usable core (env step + batch-1 torch
    # forward of both policies + sampling), bounded sample
    workers = max(1, min(cores, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else cores))
    roll = None
    try:
        secs = 6.0 if args.leg == "rollout" else 3.0
        rv, rates = cpu_sampler_all_cores(args.level, secs, workers)
        roll = {"value": rv, "unit": UNIT, "cores": workers, "kind": "port",
                "sample": f"{workers} worker processes x {secs:.0f} s: C-oracle env step + batch-1 torch-CPU forward of both policies "
                          "(actor + central critic) + MultiCategorical sampling per env step, 1 torch thread each (the reference: one "
                          "Ray rollout worker per core, train_hetero.py:212, with the slower Python env)",
                "per_worker_min_max": [min(rates), max(rates)]}
    except Exception as ex:  # noqa: BLE001
        roll = {"error": repr(ex)}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, cores_note=True),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "rollout": roll, "leg": args.leg, "gpu_launches": 0}
    if args.leg == "rollout" and roll and "value" in roll:
        line.update(value=roll["value"], ms_per_step=None, cpu_baseline=dict(roll), step_kernel_leg={"value": value, "unit": UNIT, "cores": cores},
                    e2e={"value": roll["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        line["config"] = workload_config(args, leg="rollout")
    print(json.dumps(line))


def workload_config(args, cores_note=False, leg=None):
    leg = leg or "step"
    if leg == "rollout":
        return {"workload": f"{args.arenas} arenas/GPU, 2-vs-2 level-{args.level} scripted-opponent fight, horizon "
                            f"{({1: 150, 2: 200, 3: 300}).get(args.level)}: the sampler loop -- Fight1 / Fight2 actor + central critic "
                            "forward, MultiCategorical sampling, env step, rollout buffers, GAE, action write-back (BASELINE configs[1])",
                "arenas_per_gpu": args.arenas, "level": args.level, "agent_mode": "fight", "leg": "rollout",
                "cache": "inputs larger than L2: every tick reads / writes its own slice of the ~130 MB rollout buffers, the env state "
                         "and the 5 MB of packed weights"}
    return {"workload": f"{args.arenas} arenas/GPU, 2-vs-2 level-{args.level} scripted-opponent fight, "
                        f"horizon {({1: 150, 2: 200, 3: 300}).get(args.level)}, uniform random MultiDiscrete actions, "
                        "auto-reset",
            "arenas_per_gpu": args.arenas, "level": args.level, "agent_mode": "fight", "leg": "step",
            "cache": "L2 flushed (256 MiB write) between timed steps"}


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Polls NVML (SM clock, max SM clock, power, throttle reasons) every ~5 ms on a thread."""
    BITS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
            "hw_power_brake": 0x80}

    def __init__(self, index: int):
        self.samples, self.ok, self._stop = [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _poll(self):
        nv = self.nv
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.perf_counter(), sm, rs, pw))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.004)

    def summary(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self._stop = True
        rows = [r for r in self.samples if t0 <= r[0] <= t1]
        scope = "timed region"
        if not rows:
            rows, scope = self.samples, "whole run (timed region shorter than the sampling period)"
        sm = sorted(r[1] for r in rows)
        reasons = set()
        for r in rows:
            for name, bit in self.BITS.items():
                if r[2] & bit:
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(reasons),
                "samples": len(rows), "scope": scope, "power_w_max": max((r[3] for r in rows), default=None)}


def pin_rank_to_local_cores(local: int, world: int):
    """N > 1: each rank (its Python launch loop and the host side of the e2e leg) gets its own slice of the cores NVML
    reports as local to its GPU (NUMA node / PCIe root); ranks whose GPUs share a node split that node's cores evenly."""
    if world <= 1 or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        avail = sorted(os.sched_getaffinity(0))
        pools = None
        try:
            import pynvml
            pynvml.nvmlInit()
            nw = (max(avail) // 64) + 1
            pools = []
            for g in range(world):
                words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(g), nw)
                pools.append(tuple(c for c in avail if (words[c // 64] >> (c % 64)) & 1))
            if not all(pools):
                pools = None
        except Exception:  # noqa: BLE001
            pools = None
        numa = pools is not None
        if pools is None:
            pools = [tuple(avail)] * world
        sharers = [g for g in range(world) if pools[g] == pools[local]]
        per = max(1, len(pools[local]) // len(sharers))
        k = sharers.index(local) * per
        mine = list(pools[local][k:k + per]) or list(pools[local])
        os.sched_setaffinity(0, mine)
        return {"cores": len(mine), "first": mine[0], "numa_local": numa}
    except Exception as ex:  # noqa: BLE001
        return {"error": repr(ex)}


# ------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from hhmarl_2d_b200 import VecLowLevelEnv, make_args

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n, K, W = args.arenas, args.steps, args.warmup
    affinity = pin_rank_to_local_cores(local, world)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = best_thread_count(args.level)
        v, dt = cpu_oracle_throughput(args.level, args.cpu_sample_steps, cores)
        cpu_base = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"C oracle, {cores} host threads (best of a thread-count probe; box reports "
                              f"{os.cpu_count()} logical CPUs) x {args.cpu_sample_steps} env-steps (L{args.level}, "
                              f"random actions, auto-reset), {dt:.1f} s wall"}

    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("HH_BENCH_NO_CLOCKS") else None
    env = VecLowLevelEnv(n, make_args(level=args.level), device=local, seed=0, arena_base=rank * n, autoreset=True)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    n_act = 16  # distinct action tensors cycled through
    acts = torch.stack([torch.randint(0, 13, (n_act, n, 2), device=dev, generator=g),
                        torch.randint(0, 9, (n_act, n, 2), device=dev, generator=g),
                        torch.randint(0, 2, (n_act, n, 2), device=dev, generator=g),
                        torch.randint(0, 2, (n_act, n, 2), device=dev, generator=g)], dim=-1).to(torch.int32).contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    env.reset()
    launches0 = env.launch_count

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: per-step CUDA events, L2 flush between steps
    for w in range(W):
        flush.fill_(w & 0xFF)
        env.step(acts[w % n_act])
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    barrier()
    t0 = time.perf_counter()
    launches_before = env.launch_count
    # The GPU first spins for ~10 ms so that the host loop below runs AHEAD of it for the whole timed region: every event
    # bracket then holds exactly one kernel with its launch already queued (no host launch lag inside a bracket, which at
    # N = 8 -- 8 Python loops on one host -- used to leak into the max-over-ranks).
    torch.cuda._sleep(20_000_000)
    for k in range(K):
        flush.fill_(k & 0xFF)          # evict the arena state from L2 (outside the event bracket)
        ev0[k].record()
        env.step(acts[k % n_act])
        ev1[k].record()
    barrier()
    t1 = time.perf_counter()
    gpu_launches = env.launch_count - launches_before
    clocks = sampler.summary(t0, t1) if sampler else None
    step_ms = sorted(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    total_ms = float(sum(step_ms))
    mine = torch.tensor([total_ms, step_ms[0], step_ms[len(step_ms) // 2], step_ms[-1]], dtype=torch.float64, device=dev)
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
    else:
        per_rank = [mine]
    per_rank = [[float(v) for v in t.tolist()] for t in per_rank]
    total_ms = max(r[0] for r in per_rank)
    value = world * n * K / (total_ms * 1e-3)
    step_stats = {"per_rank_us": [{"mean": 1e3 * r[0] / K, "min": 1e3 * r[1], "median": 1e3 * r[2], "max": 1e3 * r[3]} for r in per_rank],
                  "note": "CUDA-event bracket of every step kernel launch; the host loop runs ahead of the GPU (a 10 ms device-side spin "
                          "precedes the timed region), `value` uses the slowest rank's sum"}

    # ---- back-to-back (no flush; state stays in L2), one event pair around K launches
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        env.step(acts[k % n_act])
    e1.record()
    barrier()
    b2b_ms = e0.elapsed_time(e1)

    # ---- on-device sampler: Fight1/Fight2 forward (actor + central critic), MultiCategorical sampling, env step,
    #      rollout-buffer writes, GAE and the action write-back, one CUDA graph per 20-tick fragment
    rollout = None
    try:
        if args.no_rollout:
            raise RuntimeError("skipped (--no-rollout)")
        from hhmarl_2d_b200 import VecSampler, TorchPolicy
        from hhmarl_2d_b200 import models as M
        rollout = {}
        Tf = 20
        for tag, kw in (("fused_tc", dict(fused="tc")), ("fused_tc_one_group", dict(fused="tc", groups=1)),
                        ("fused_3xtf32", dict(fused="3xtf32", groups=1)), ("fused_tf32", dict(fused="tf32", groups=1)),
                        ("fp32", dict(allow_tf32=False, fused=None)), ("tf32", dict(allow_tf32=True, fused=None)),
                        ("fp32_unpacked", dict(allow_tf32=False, packed=False, fused=None))):
            torch.manual_seed(rank)
            m1, m2 = M.build_policy_pair("fight")
            m1.to(dev); m2.to(dev)
            env_r = VecLowLevelEnv(n, make_args(level=args.level), device=local, seed=1, arena_base=rank * n, autoreset=True)
            smp = VecSampler(env_r, TorchPolicy(m1, 1), TorchPolicy(m2, 2), fragment_len=Tf, use_cuda_graph=True, **kw)
            for _ in range(3):
                smp.collect()
            R = max(2, K // Tf)
            barrier()
            r_sampler = ClockSampler(local) if (rank == 0 and tag == "fused_tc" and not os.environ.get("HH_BENCH_NO_CLOCKS")) else None
            tr0 = time.perf_counter()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            for _ in range(R):
                smp.collect()
            r1.record()
            barrier()
            if r_sampler is not None:       # SM clock / throttle reasons DURING the rollout's timed region (the --leg rollout line's `clocks`)
                rollout["clocks"] = r_sampler.summary(tr0, time.perf_counter())
                rollout["launches_per_fragment"] = 4 + 3 * Tf * getattr(smp, "groups", 1)
            rt = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(rt, op=dist.ReduceOp.MAX)
            rollout[tag] = {"value": world * n * Tf * R / (float(rt.item()) * 1e-3), "unit": UNIT,
                            "ms_per_tick": float(rt.item()) / (R * Tf)}
            if tag == "fused_tc":
                rollout["fragment_bytes"] = int(sum(v.numel() * v.element_size() for v in smp.buf.values()))
                # (i) the same loop end to end: every fragment's batch is copied to pinned host memory inside the timed region
                #     (what a host-side learner / RLlib's train batch would read); the sampler has no host INPUT
                #     VecSampler(n_buffers=2).collect_host(): two buffer sets, fragment k travels D2H while k + 1 is sampled
                env_h = VecLowLevelEnv(n, make_args(level=args.level), device=local, seed=1, arena_base=rank * n, autoreset=True)
                smp_h = VecSampler(env_h, TorchPolicy(m1, 1), TorchPolicy(m2, 2), fragment_len=Tf, use_cuda_graph=True,
                                   n_buffers=2, **kw)
                for _ in range(6):
                    smp_h.collect_host()
                torch.cuda.synchronize()
                barrier()
                th0 = time.perf_counter()
                for _ in range(R):
                    _, h_ev = smp_h.collect_host()
                h_ev.synchronize()
                torch.cuda.synchronize()
                th1 = time.perf_counter()
                et = torch.tensor([th1 - th0], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(et, op=dist.ReduceOp.MAX)
                rollout["e2e"] = {"value": world * n * Tf * R / float(et.item()), "unit": UNIT, "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": rollout["fragment_bytes"] // Tf,
                                  "api": "VecSampler(n_buffers=2).collect_host(): every fragment's whole batch copied into pinned host "
                                         "memory inside the timed region, the copy of fragment k under the sampling of fragment k + 1"}
                del smp_h, env_h
                # (ii) the dominant kernel alone: the tcgen05 policy forward on the recorded central observations of the
                #      fragment (a different tick's rows every launch), CUDA events around R2 launches
                fu, bufs = smp.packed, smp.buf
                outs = (torch.empty((n, 26), device=dev), torch.empty((n,), device=dev), torch.empty((n, 24), device=dev),
                        torch.empty((n,), device=dev))
                for t in range(3):
                    fu.forward(bufs["flat1"][t], bufs["flat2"][t], out=outs)
                R2 = 40
                k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                k0.record()
                for t in range(R2):
                    fu.forward(bufs["flat1"][t % Tf], bufs["flat2"][t % Tf], out=outs)
                k1.record()
                torch.cuda.synchronize()
                kern_us = k0.elapsed_time(k1) * 1e3 / R2
                macs_useful = macs_issued = 0
                for p_ in range(2):
                    for kk in range(2):
                        w1 = fu._rm["w1"][p_][kk]
                        att_n = fu.att[kk][1] if fu.fight else 0
                        n_out = fu.packed.Wact[p_].shape[1] if kk == 0 else 1
                        macs_useful += int((w1 != 0).sum()) + att_n * att_n + 500 * 500 + 500 * n_out
                        k1s = (fu.packed.W1[p_].shape[0] + 15) // 16
                        att_iss = 0
                        if fu.fight:
                            from hhmarl_2d_b200.fused_forward import _att_image_geometry
                            _, att_ks, att_nn = _att_image_geometry(fu.att[kk][0], fu.att[kk][1])
                            att_iss = att_ks * 16 * att_nn
                        macs_issued += k1s * 16 * 512 + att_iss + 512 * 512 + 512 * 32
                from hhmarl_2d_b200 import _native as _nat
                tc_mode = int(_nat.lib().hh_policy_tc_mode())         # 2: 128-row tiles (default), 0: 64-row, 1: CTA pairs
                tile_rows = 64 if tc_mode == 0 else 128
                kname = {0: "policy_forward_tc_kernel", 1: "policy_forward_pair_kernel", 2: "policy_forward_m128_kernel"}[tc_mode]
                rows_pad = (n + tile_rows - 1) // tile_rows * tile_rows
                peak_tf, peak_src2 = 1590.0, "fallback (B200_PROFILING.md)"
                try:
                    peak_tf = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
                    peak_src2 = "measured (MEASURED_PEAKS.json bf16_tflops, cuBLAS burst; kind::f16 and bf16 MMAs run at the same rate)"
                except Exception:  # noqa: BLE001
                    pass
                issued = 3 * 2 * macs_issued * rows_pad / (kern_us * 1e-6) / 1e12
                ptraffic, ptensor = None, None
                try:
                    pj = json.load(open(os.path.join(ROOT, "profiles", "policy_kernel_traffic.json")))
                    ptraffic, ptensor = pj.get("dram_bytes_per_launch"), pj.get("tensor_pipe_active_pct")
                except Exception:  # noqa: BLE001
                    pass
                rollout["roofline"] = {
                    "bound": "tensor", "achieved": issued, "peak": peak_tf, "unit": "TFLOP/s", "frac": issued / peak_tf,
                    "traffic": ptraffic, "tensor_pipe_active_pct_ncu": ptensor, "peak_source": peak_src2, "kernel": "hh::tc::" + kname,
                    "tile_rows": tile_rows, "kernel_us": kern_us,
                    "useful_tflops": 2 * macs_useful * n / (kern_us * 1e-6) / 1e12,
                    "useful_flop_per_launch": 2 * macs_useful * n, "issued_flop_per_launch": 3 * 2 * macs_issued * rows_pad,
                    "share_of_rollout_tick": kern_us * 1e-3 / rollout[tag]["ms_per_tick"],
                    "note": "fp32-equivalent forward = 3 kind::f16 MMAs per product on zero-padded tiles (K to 16, N to 256 / 104 / "
                            "112 / 160 / 32): `achieved` counts the MMA work issued, `useful_tflops` the reference's own multiply-adds.  "
                            + ("128-row tiles (M = 128 MMAs at the full tensor rate; A_hi in shared memory, A_lo in tensor memory); the "
                               "layer chain of one tile is serial (MMA -> tanh epilogue -> next layer's operand) and one CTA per SM "
                               "fits, so the tensor pipe idles during the exposed epilogues and the 256 tiles take two rounds on 148 "
                               "SMs (0.86 occupancy of the second round): see DESIGN.md section 4 for the per-tile timeline"
                               if tc_mode == 2 else
                               "An M = 64 tile runs the tensor pipe at half the M = 128 rate (profiles/r2a_tcgen05_probe.txt), so 0.5 is "
                               "the ceiling of `frac` for this tile shape")}
            del smp, env_r
        rollout["fragment_len"] = Tf
        if cpu_base is not None:   # N = 1, rank 0: the reference-style rollout worker on one host core, bounded sample
            try:
                v1 = cpu_sampler_equivalent(args.level)
                rollout["cpu_sampler_equivalent"] = {
                    "value_per_worker": v1, "unit": UNIT, "cores": 1,
                    "sample": "4 s of one worker: C-oracle env step + batch-1 torch-CPU forward of both policies (actor + "
                              "central critic) + MultiCategorical sampling, 1 torch thread; the reference runs one such "
                              "worker per core (train_hetero.py:212) with the Python env instead of the C oracle"}
            except Exception as ex:  # noqa: BLE001
                rollout["cpu_sampler_equivalent"] = {"error": repr(ex)}
        rollout["note"] = ("both policies' actor + central critic + Gumbel-max sampling + env step + GAE + action "
                           "write-back, one CUDA graph per 20-tick fragment; random-init weights.  'fused_tc' (the sampler's "
                           "default): csrc/hh_policy_tc.cu, tcgen05 / TMEM forward, fp32-equivalent (fp16 hi/lo operand split), the "
                           "batch advanced as 8 parts on their own streams inside the graph (a part steps while the other parts' "
                           "forwards run; 'fused_tc_one_group' = the same kernels on one stream); "
                           "'fused_3xtf32' / 'fused_tf32': the mma.sync forward kernel csrc/hh_policy.cu (one launch "
                           "per tick, 3xTF32 = fp32-equivalent / plain TF32 tensor-core products); 'fp32' / 'tf32': packed "
                           "cuBLAS GEMMs (fused_forward.PackedPolicyPair); 'fp32_unpacked': per-layer forward of models.py")
    except Exception as ex:  # noqa: BLE001
        rollout = {"error": repr(ex)}

    # ---- BASELINE config 5: 3-vs-3 hierarchical commander rollout (HighLevelEnv), random commander actions,
    #      every aircraft's frozen low-level policy batched across arenas (random-init weights)
    hier = None
    try:
        if args.no_hier:
            raise RuntimeError("skipped (--no-hier)")
        from hhmarl_2d_b200.env_hier import VecHighLevelEnv
        henv = VecHighLevelEnv(n, device=local, seed=2, arena_base=rank * n, autoreset=True)
        henv.reset()
        gh = torch.Generator(device=dev)
        gh.manual_seed(77 + rank)
        cmd = torch.randint(0, 3, (8, n, 3), device=dev, generator=gh).to(torch.int32)
        tick_sum = torch.zeros((), dtype=torch.int64, device=dev)
        for w in range(5):            # two eager steps, the capture of the step's CUDA graph, two replays
            henv.step(cmd[w])
            tick_sum += henv.substeps.sum()      # (also loads the reduction kernel outside the timed region)
        barrier()
        tick_sum.zero_()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        HS = 5
        hev = [torch.cuda.Event(enable_timing=True) for _ in range(HS + 1)]
        hev[0].record()
        for k in range(HS):
            henv.step(cmd[(5 + k) % 8])
            tick_sum += henv.substeps.sum()
            hev[k + 1].record()
        h1.record()
        barrier()
        ticks = int(tick_sum.item())
        hier_step_ms = [hev[k].elapsed_time(hev[k + 1]) for k in range(HS)]
        ht = torch.tensor([h0.elapsed_time(h1)], dtype=torch.float64, device=dev)
        tk = torch.tensor([float(ticks)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ht, op=dist.ReduceOp.MAX)
            dist.all_reduce(tk)
        hier = {"commander_steps_per_s": world * n * HS / (float(ht.item()) * 1e-3),
                "sim_ticks_per_s": float(tk.item()) / (float(ht.item()) * 1e-3), "commander_steps": HS,
                "mean_substeps": float(tk.item()) / (world * n * HS), "arenas_per_gpu": n, "step_ms": hier_step_ms,
                "note": "HighLevelEnv 3-vs-3, 16 masked sub-steps x (2 staged env launches + 2 launches of csrc/hh_policy_tc.cu: the frozen "
                        "fight / escape actors of all six aircraft as gathered chains on tcgen05, fp32-equivalent, argmax in the epilogue; "
                        "row lists built on the device, no host synchronisation inside a commander step: replayed as one CUDA graph)"}
        from hhmarl_2d_b200.env_hier import CommanderSampler
        from hhmarl_2d_b200 import models as MM
        cs = CommanderSampler(henv, MM.CommanderGru().to(dev), fragment_len=4)
        cs.collect()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        cs.collect()
        c1.record()
        barrier()
        ct = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ct, op=dist.ReduceOp.MAX)
        hier["commander_rollout_steps_per_s"] = world * n * 4 / (float(ct.item()) * 1e-3)
        del henv, cs
    except Exception as ex:  # noqa: BLE001
        hier = {"error": repr(ex)}

    # ---- BASELINE configs 3/4: level-5 self-play (frozen fight / escape policy sets as opponents, one set per arena and
    #      episode): hh_step_begin -> opponents' actors -> hh_step_finish, random agent actions, random-init weights
    l5 = None
    try:
        if args.no_l5:
            raise RuntimeError("skipped (--no-l5)")
        l5 = {}
        # BASELINE config 3 = 32 768 arenas on ONE GPU; config 4 = 8 192 per GPU on 8 (65 536 in total)
        n5 = args.l5_arenas or (32768 if world == 1 else 8192)
        g5 = torch.Generator(device=dev)
        g5.manual_seed(4321 + rank)
        acts5 = torch.stack([torch.randint(0, 13, (4, n5, 2), device=dev, generator=g5), torch.randint(0, 9, (4, n5, 2), device=dev, generator=g5),
                             torch.randint(0, 2, (4, n5, 2), device=dev, generator=g5), torch.randint(0, 2, (4, n5, 2), device=dev, generator=g5)],
                            dim=-1).to(torch.int32).contiguous()
        for tag, fused in (("fused_actors", True), ("torch_actors", False)):
            env5 = VecLowLevelEnv(n5, make_args(level=5), device=local, seed=3, arena_base=rank * n5, autoreset=True,
                                  allow_standin_opponents=True)   # random-init frozen actors: no trained weights exist here
            env5.fused_opponents = fused
            env5.reset()
            for w in range(5):
                env5.step(acts5[w % 4])
            barrier()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            S5 = 40 if fused else 8
            q0.record()
            for k in range(S5):
                env5.step(acts5[k % 4])
            q1.record()
            barrier()
            qt = torch.tensor([q0.elapsed_time(q1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(qt, op=dist.ReduceOp.MAX)
            l5[tag] = {"value": world * n5 * S5 / (float(qt.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(qt.item()) / S5}
            del env5
        l5["arenas_per_gpu"] = n5
        l5["config"] = "BASELINE configs[2]: 32 768 arenas, level 5, 1 GPU" if world == 1 and n5 == 32768 else \
            f"BASELINE configs[3] layout: {n5} arenas per GPU x {world} GPUs, level 5"
        l5["note"] = ("level 5 (horizon 400): split step around the opponents' frozen actors (3 policy sets x 2 aircraft types); "
                      "'fused_actors' = csrc/hh_policy_tc.cu chains (one tcgen05 launch, fp32-equivalent, argmax in the epilogue), "
                      "'torch_actors' = per-set gather + per-layer cuBLAS forward")
        del acts5
    except Exception as ex:  # noqa: BLE001
        l5 = {"error": repr(ex)}

    # ---- BASELINE config 4 (and its 1-GPU slice): level-5 rollout fragment -> PPO update with the gradient all-reduce over
    #      NCCL (the ONE collective of the path), 8 192 arenas per GPU
    ppo = None
    try:
        if args.no_ppo:
            raise RuntimeError("skipped (--no-ppo)")
        from hhmarl_2d_b200 import VecSampler, TorchPolicy, PPOLearner
        from hhmarl_2d_b200 import models as M
        torch.manual_seed(0)                       # identical initial weights on every rank
        m1, m2 = M.build_policy_pair("fight")
        m1.to(dev); m2.to(dev)
        envp = VecLowLevelEnv(n, make_args(level=5), device=local, seed=5, arena_base=rank * n, autoreset=True,
                              allow_standin_opponents=True)
        Tp = 20
        smp = VecSampler(envp, TorchPolicy(m1, 1), TorchPolicy(m2, 2), fragment_len=Tp, use_cuda_graph=True)
        learner = PPOLearner(m1, m2, num_sgd_iter=1, sgd_minibatch_size=8192)
        learner.time_allreduce = True
        for _ in range(3):           # (the remainder minibatch comes once per update: its graph is captured in the third)
            learner.update(smp.collect())
            smp.refresh_policy()
        IT = 3
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * IT + 1)]
        barrier()
        ev[0].record()
        for it in range(IT):
            batch = smp.collect()
            ev[2 * it + 1].record()
            st = learner.update(batch)
            smp.refresh_policy()
            ev[2 * it + 2].record()
        barrier()
        t_s = sum(ev[2 * it].elapsed_time(ev[2 * it + 1]) for it in range(IT))
        t_l = sum(ev[2 * it + 1].elapsed_time(ev[2 * it + 2]) for it in range(IT))
        tt = torch.tensor([t_s, t_l, t_s + t_l], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_s, t_l, t_all = (float(v) for v in tt.tolist())
        ar = learner.allreduce_us()
        ppo = {"sample_and_learn": {"value": world * n * Tp * IT / (t_all * 1e-3), "unit": UNIT},
               "rollout": {"value": world * n * Tp * IT / (t_s * 1e-3), "unit": UNIT, "ms_per_fragment": t_s / IT},
               "learner_ms_per_update": t_l / IT, "minibatches_per_update": st["minibatches"], "sgd_minibatch_size": 8192,
               "num_sgd_iter": 1, "fragment_len": Tp, "arenas_per_gpu": n, "level": 5,
               "allreduce": {"bytes": learner.grad_bytes, "us_per_minibatch": ar, "backend": "nccl" if world > 1 else None,
                             "world": world},
               "note": "level-5 self-play rollout (frozen stand-in opponents) + PPOLearner.update (both policies, flat "
                       "parameter / gradient buffers, fused Adam), one all-reduce of the flat gradient per minibatch; "
                       "num_sgd_iter 1 here to bound the bench (train_hetero.py uses the RLlib default 30)"}
        try:      # the same update with the learner's GEMMs as TF32 tensor-core products (PPOLearner(matmul_tf32=True); not the default)
            m1t, m2t = M.build_policy_pair("fight")
            m1t.to(dev); m2t.to(dev)
            lt = PPOLearner(m1t, m2t, num_sgd_iter=1, sgd_minibatch_size=8192, matmul_tf32=True)
            bt = smp.collect()
            for _ in range(3):
                lt.update(bt)
            barrier()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            for _ in range(IT):
                lt.update(bt)
            q1.record()
            barrier()
            qt = torch.tensor([q0.elapsed_time(q1) / IT], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(qt, op=dist.ReduceOp.MAX)
            ppo["learner_ms_per_update_matmul_tf32"] = float(qt.item())
            del lt
        except Exception as ex:  # noqa: BLE001
            ppo["learner_ms_per_update_matmul_tf32"] = repr(ex)
        del smp, envp, learner
    except Exception as ex:  # noqa: BLE001
        ppo = {"error": repr(ex)}

    # ---- end to end through the host entry point of the C ABI
    acts_host = acts.cpu().numpy()
    pin_act, *pin_out = env.host_buffers()       # the handle's pinned slab: the driver writes actions in place
    outs = tuple(pin_out)
    for w in range(max(3, W // 4)):
        pin_act[...] = acts_host[w % n_act]
        env.step_host(pin_act, out=outs)
    barrier()
    th0 = time.perf_counter()
    for k in range(K):
        pin_act[...] = acts_host[k % n_act]      # the caller producing this step's actions (262 KB host write)
        env.step_host(pin_act, out=outs)
    torch.cuda.synchronize()
    th1 = time.perf_counter()
    e2e_t = torch.tensor([th1 - th0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * K / float(e2e_t.item())
    # the other host modes of the same entry point (hh_set_host_mode), for comparison
    e2e_alt = {}
    default_mode = os.environ.get("HH_HOST_MODE", "zerocopy")
    for alt_mode in [m for m in ("pipelined", "zerocopy", "staged") if m != default_mode]:
        try:
            env.set_host_mode(alt_mode)
            for w in range(3):
                env.step_host(pin_act, out=outs)
            barrier()
            ta0 = time.perf_counter()
            for k in range(K):
                pin_act[...] = acts_host[k % n_act]
                env.step_host(pin_act, out=outs)
            torch.cuda.synchronize()
            ta1 = time.perf_counter()
            alt_t = torch.tensor([ta1 - ta0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(alt_t, op=dist.ReduceOp.MAX)
            e2e_alt[alt_mode] = {"value": world * n * K / float(alt_t.item()), "unit": UNIT}
        except Exception as ex:  # noqa: BLE001
            e2e_alt[alt_mode] = {"error": repr(ex)}
    try:
        env.set_host_mode(default_mode)
    except Exception:  # noqa: BLE001
        pass
    d1, d2 = env.obs_dim
    h2d, d2h = n * 8 * 4, n * ((d1 + d2 + 2) * 4 + 1)

    # ---- the same workload through the send / poll halves of the host call (hh_step_host_begin / _end), the batch split
    #      over two handles that are kept in flight together: one half's PCIe traffic and the caller's preparation of
    #      the next actions overlap the other half's kernel.  Every arena still takes one step per round, with its
    #      actions coming from and its results going to host memory inside the timed region.
    e2e_pipe = None
    try:
        half = n // 2
        envs2 = [VecLowLevelEnv(half, make_args(level=args.level), device=local, seed=0, arena_base=rank * n + i * half,
                                autoreset=True) for i in range(2)]
        bufs2 = [e.host_buffers() for e in envs2]
        for e in envs2:
            e.reset_host()

        def one_round(k):
            src = acts_host[k % n_act]
            for i, e in enumerate(envs2):
                bufs2[i][0][...] = src[i * half:(i + 1) * half]
                e.send_actions_host(bufs2[i][0])
            for i, e in enumerate(envs2):
                e.poll_host(out=tuple(bufs2[i][1:]))

        for w in range(max(3, W // 4)):
            one_round(w)
        barrier()
        tp0 = time.perf_counter()
        for k in range(K):
            one_round(k)
        torch.cuda.synchronize()
        tp1 = time.perf_counter()
        pipe_t = torch.tensor([tp1 - tp0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(pipe_t, op=dist.ReduceOp.MAX)
        e2e_pipe = {"value": world * n * K / float(pipe_t.item()), "unit": UNIT,
                    "api": "send_actions_host / poll_host (hh_step_host_begin / _end) on two half-batch handles in flight together"}
        del envs2, bufs2
    except Exception as ex:  # noqa: BLE001
        e2e_pipe = {"error": repr(ex)}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
        if os.path.exists(peaks_path):
            try:
                peak = float(json.load(open(peaks_path))["hbm_gbs"])
                peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
            except Exception:  # noqa: BLE001
                pass
        kern_ms = total_ms / K                      # one kernel per step: event bracket == the launch
        achieved = ALGO_BYTES_PER_ENV_STEP * n / (kern_ms * 1e-3) / 1e9
        traffic, attainable = None, None
        tp = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("dram_bytes_per_launch")
                if tj.get("issue_active_pct") and n == 8192:
                    # the bound this kernel can actually approach: its 4.44 M warp instructions on 148 x 4 issue slots
                    w = float(tj["warp_instructions_per_launch"])
                    attainable = {"bound": "issue slots", "achieved_pct_of_peak": tj["issue_active_pct"],
                                  "fp64_pipe_active_pct": tj.get("fp64_pipe_active_pct"),
                                  "warp_instructions_per_launch": w,
                                  "floor_us_at_full_issue": w / (148 * 4) / 1.965e3,
                                  "stall_no_instruction_per_issue": tj.get("stall_no_instruction_per_issue"),
                                  "source": tj.get("counters_source"),
                                  "note": "measured ncu counters: the step is a chain of ten barrier-separated stages with one or two "
                                          "warps per scheduler; instruction fetch (no_instruction) and fixed FP64 latencies (wait) leave "
                                          "three of four issue slots empty -- it is latency-bound, neither DRAM- nor FP64-throughput-bound"}
            except Exception:  # noqa: BLE001
                pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(args),
                "agent_steps_per_s": 2 * value,
                "back_to_back": {"value": world * n * K / (b2b_ms * 1e-3), "unit": UNIT,
                                 "note": "no L2 flush, K launches under one event pair (rank 0)"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "hh_step_host on the pinned host buffers of hh_host_buffers (C ABI), host mode "
                               + os.environ.get("HH_HOST_MODE", "zerocopy") + " (pipelined = two half-batch launches in one CUDA graph, the first half's "
                               "observations travel D2H under the second half's kernel, actions / rewards / done flags through the pinned "
                               "slab; zerocopy = the kernel reads / writes the pinned slab over PCIe; staged = 1 H2D + launch + 1 D2H)",
                        "other_host_modes": e2e_alt, "send_poll_two_handles": e2e_pipe},
                "rollout": rollout,
                "hier": hier,
                "level5": l5,
                "ppo": ppo,
                "step_time": step_stats,
                "cpu_affinity": affinity,
                "leg": args.leg,
                "gpu_launches": int(gpu_launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * n,
                             "kernel": {"quad": "hh::step_kernel<3,0>", "cta": "hh::step_kernel_cta<3,0>"}.get(
                                 os.environ.get("HH_STEP_IMPL", "v4"), "hh::step_kernel_v4<3,0>"),
                             "kernel_ms": kern_ms, "attainable": attainable,
                             "note": "bound by the dependent FP64 instruction chain of one CTA's step phases, not by DRAM: see DESIGN.md section 4"},
                "clocks": clocks}
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        if args.leg == "rollout" and isinstance(rollout, dict) and "fused_tc" in rollout:
            # `value` = the whole sampler loop (north_star's "rollout loop"); the step-kernel numbers move to step_kernel_leg
            line["step_kernel_leg"] = {k: line[k] for k in ("value", "ms_per_step", "e2e", "roofline", "gpu_launches", "config")}
            r = rollout["fused_tc"]
            line.update(value=r["value"], ms_per_step=r["ms_per_tick"], config=workload_config(args, leg="rollout"),
                        dtype="f64 env + fp32-equivalent policy (fp16 hi/lo split on tcgen05)",
                        e2e=rollout.get("e2e"), roofline=rollout.get("roofline"),
                        gpu_launches=int(rollout.get("launches_per_fragment", 4 * rollout["fragment_len"])) * max(2, K // 20))
            if rollout.get("clocks"):
                line["clocks"] = rollout["clocks"]
            if cpu_base is not None and isinstance(rollout.get("cpu_sampler_equivalent"), dict) and "value_per_worker" in rollout["cpu_sampler_equivalent"]:
                c = rollout["cpu_sampler_equivalent"]
                line["cpu_baseline"] = {"value": c["value_per_worker"], "unit": UNIT, "cores": 1, "kind": "port", "sample": c["sample"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
