#!/bin/bash
# profiles/run_r2q.sh -- round-2 evidence for the tcgen05 policy forward: pair-variant test, ncu launch list of the rollout leg,
# ncu --set full capture of the kernel, refreshed bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampler.py -m gpu -x -q -k "pair_kernel or pack_image or fused_policy_forward or actor_chains" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r2q.log
timeout 600 python bench.py --leg rollout --steps 100 --no-hier --no-l5 --no-ppo > gpurun_out/bench_r2q_rollout.json 2> gpurun_out/bench_r2q.err
timeout 900 python bench.py > gpurun_out/bench_r2q.json 2>> gpurun_out/bench_r2q.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches_r2q_rollout.csv \
  python bench.py --leg rollout --steps 40 --warmup 3 --no-cpu-baseline --no-hier --no-l5 --no-ppo > gpurun_out/ncu_r2q_list.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:policy_forward_tc -s 6 -c 2 -o gpurun_out/prof_policy_r2q -f \
  python profiles/policy_forward_probe.py 8192 > gpurun_out/ncu_r2q_full.log 2>&1
tail -3 gpurun_out/ncu_r2q_full.log
python - <<PY
import json
for f in ("bench_r2q.json", "bench_r2q_rollout.json"):
    d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    r = d["rollout"]
    print(f, "value", round(d["value"] / 1e6, 2), "M; rollout fused_tc", round(r["fused_tc"]["value"] / 1e6, 2), "M", r["fused_tc"]["ms_per_tick"],
          "kernel_us", r["roofline"]["kernel_us"], "frac", round(r["roofline"]["frac"], 3), "useful", round(r["roofline"]["useful_tflops"], 1))
    if "level5" in d and isinstance(d["level5"], dict) and "fused_actors" in d["level5"]:
        print("  l5", round(d["level5"]["fused_actors"]["value"] / 1e6, 1), "M; hier", round(d["hier"]["commander_steps_per_s"] / 1e3), "k; ppo", json.dumps(d["ppo"])[:300])
PY
