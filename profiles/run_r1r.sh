#!/bin/bash
# profiles/run_r1r.sh -- last pass of round 1: full GPU tests, smoke, bench
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r1r.log
echo "== smoke"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench (default)"
timeout 900 python bench.py 2> gpurun_out/bench_r1r.err | tail -1 > gpurun_out/bench_r1r.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r1r.json'))
print("value", d["value"] / 1e6, "us", d["ms_per_step"] * 1e3, "e2e", d["e2e"]["value"] / 1e6, "frac", d["roofline"]["frac"])
print({k: round(v["value"] / 1e6, 2) for k, v in d["rollout"].items() if isinstance(v, dict)})
print("hier", d["hier"].get("commander_steps_per_s"), d["hier"].get("sim_ticks_per_s"), "l5", d["level5"].get("fused_actors"))
PY
tail -2 gpurun_out/bench_r1r.err
