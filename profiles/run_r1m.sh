#!/bin/bash
# profiles/run_r1m.sh -- fused policy-forward kernel (csrc/hh_policy.cu): numerics, sampler tests, rollout bench
mkdir -p gpurun_out
echo "== sampler tests"
timeout 900 python -m pytest tests/test_gpu_sampler.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_sampler_r1m.log
echo "== bench"
timeout 600 python bench.py --no-cpu-baseline --no-hier 2> gpurun_out/bench_r1m.err | tail -1 > gpurun_out/bench_r1m.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r1m.json'))
print(json.dumps(d["rollout"], indent=1))
PY
tail -5 gpurun_out/bench_r1m.err
