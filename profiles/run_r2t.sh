#!/bin/bash
# profiles/run_r2t.sh -- full GPU suite + default bench after the staged hierarchical kernels and the device-built row lists
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r2t.log
timeout 900 python bench.py > gpurun_out/bench_r2t.json 2> gpurun_out/bench_r2t.err; tail -c 800 gpurun_out/bench_r2t.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2t.json").read().strip().splitlines()[-1])
print("value", round(d["value"] / 1e6, 1), "e2e", round(d["e2e"]["value"] / 1e6, 1), "rollout", round(d["rollout"]["fused_tc"]["value"] / 1e6, 1))
print("hier", json.dumps(d["hier"])[:400])
print("l5", json.dumps(d["level5"])[:300])
print("ppo", json.dumps(d["ppo"])[:300])
PY
