#!/bin/bash
# profiles/run_r2_final.sh -- end-of-round validation on one B200: smoke, the whole GPU test suite, both bench arms (default leg and
# --leg rollout)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/smoke_r2_final.log
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r2_final.log
timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
timeout 600 python bench.py --leg rollout --steps 100 --no-hier --no-l5 --no-ppo > gpurun_out/bench_r2_final_rollout.json 2>> gpurun_out/bench_r2_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_r2_final_reference.json 2>> gpurun_out/bench_r2_final.err
timeout 600 python bench.py --impl reference --leg rollout --steps 5 --warmup 3 > gpurun_out/bench_r2_final_reference_rollout.json 2>> gpurun_out/bench_r2_final.err
tail -c 400 gpurun_out/bench_r2_final.err
python - <<PY
import json
for f in ("bench_r2_final.json", "bench_r2_final_rollout.json", "bench_r2_final_reference.json", "bench_r2_final_reference_rollout.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "unreadable", ex); continue
    print(f, d.get("impl", "own"), "leg", d.get("leg"), "value %.3g" % d["value"], "e2e %.3g" % d["e2e"]["value"], "ms_per_step", d.get("ms_per_step"))
    r = d.get("rollout")
    if isinstance(r, dict) and "fused_tc" in r:
        print("  rollout", {k: round(v["value"] / 1e6, 1) for k, v in r.items() if isinstance(v, dict) and "ms_per_tick" in v}, "e2e", round(r["e2e"]["value"] / 1e6, 1),
              "kernel_us", round(r["roofline"]["kernel_us"], 1), "frac", round(r["roofline"]["frac"], 3))
    if isinstance(d.get("level5"), dict) and "fused_actors" in d["level5"]:
        print("  l5", round(d["level5"]["fused_actors"]["value"] / 1e6, 1), "M; hier", round(d["hier"]["commander_steps_per_s"] / 1e3), "k; ppo", json.dumps(d["ppo"])[:260])
PY
