#!/bin/bash
# profiles/sweep_impl.sh -- quad (v2) vs cta (v3) fused step at several arena counts (run under gpurun)
for impl in quad cta; do
  for n in 8192 32768 131072; do
    HH_STEP_IMPL=$impl timeout 200 python bench.py --arenas $n --steps 100 --warmup 10 --no-cpu-baseline --no-rollout --no-hier 2>&1 | tail -1 > /tmp/l.json
    python - "$impl" "$n" <<'PY'
import json, sys
d = json.load(open('/tmp/l.json'))
print(sys.argv[1], sys.argv[2], "arenas:", round(d["value"] / 1e6, 1), "M env-steps/s,", round(d["ms_per_step"] * 1000, 1), "us/step")
PY
  done
done
