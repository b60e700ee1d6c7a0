#!/bin/bash
# profiles/sweep_cta_arenas.sh -- v3 fused step with 8 / 16 / 32 arenas per CTA (run under gpurun)
for a in 8 16 32; do
  for n in 8192 32768; do
    HH_STEP_IMPL=cta HH_LIB_PATH=$PWD/build/lib_a$a.so timeout 200 python bench.py --arenas $n --steps 200 --warmup 20 --no-cpu-baseline --no-rollout --no-hier 2>&1 | tail -1 > /tmp/l.json
    python - "$a" "$n" <<'PY'
import json, sys
d = json.load(open('/tmp/l.json'))
print("arenas/CTA", sys.argv[1], ",", sys.argv[2], "arenas:", round(d["value"] / 1e6, 1), "M env-steps/s,", round(d["ms_per_step"] * 1000, 1), "us/step")
PY
  done
done
HH_STEP_IMPL=cta HH_LIB_PATH=$PWD/build/lib_a8.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "oracle_many or full_size" 2>&1 | tail -1
HH_STEP_IMPL=cta HH_LIB_PATH=$PWD/build/lib_a16.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "oracle_many or full_size" 2>&1 | tail -1
