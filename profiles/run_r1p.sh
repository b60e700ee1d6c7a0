#!/bin/bash
# profiles/run_r1p.sh -- fused actors for the level-4/5 opponents and the hierarchical env: full tests + bench
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r1p.log
echo "== bench (default)"
timeout 900 python bench.py 2> gpurun_out/bench_r1p.err | tail -1 > gpurun_out/bench_r1p.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r1p.json'))
print("value", d["value"] / 1e6, "e2e", d["e2e"]["value"] / 1e6)
print(json.dumps({k: d[k] for k in ("rollout", "hier", "level5")}, indent=1)[:3000])
PY
tail -3 gpurun_out/bench_r1p.err
