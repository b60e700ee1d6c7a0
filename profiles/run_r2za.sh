#!/bin/bash
# profiles/run_r2za.sh -- evidence for the 128-row tcgen05 policy forward (the default): GPU tests of the policy path, rollout leg,
# default bench, ncu launch list of the rollout leg, ncu --set full capture of the kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampler.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r2za.log
timeout 600 python bench.py --leg rollout --steps 100 --no-hier --no-l5 --no-ppo > gpurun_out/bench_r2za_rollout.json 2> gpurun_out/bench_r2za.err
timeout 900 python bench.py > gpurun_out/bench_r2za.json 2>> gpurun_out/bench_r2za.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches_r2za_rollout.csv \
  python bench.py --leg rollout --steps 40 --warmup 3 --no-cpu-baseline --no-hier --no-l5 --no-ppo > gpurun_out/ncu_r2za_list.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:policy_forward_m128 -s 6 -c 2 -o gpurun_out/prof_policy_r2za -f \
  python profiles/policy_forward_probe.py 8192 > gpurun_out/ncu_r2za_full.log 2>&1
tail -3 gpurun_out/ncu_r2za_full.log
python - <<PY
import json
for f in ("bench_r2za.json", "bench_r2za_rollout.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "unreadable", ex); continue
    r = d["rollout"]
    print(f, "value", round(d["value"] / 1e6, 2), "M; e2e", round(d["e2e"]["value"] / 1e6, 2), "M; rollout fused_tc", round(r["fused_tc"]["value"] / 1e6, 2), "M", r["fused_tc"]["ms_per_tick"],
          "kernel_us", r["roofline"]["kernel_us"], "frac", round(r["roofline"]["frac"], 3), "useful", round(r["roofline"]["useful_tflops"], 1))
    if "level5" in d and isinstance(d["level5"], dict) and "fused_actors" in d["level5"]:
        print("  l5", round(d["level5"]["fused_actors"]["value"] / 1e6, 1), "M; hier", round(d["hier"]["commander_steps_per_s"] / 1e3), "k; ppo", json.dumps(d["ppo"])[:300])
PY
