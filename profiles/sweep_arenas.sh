#!/bin/bash
# profiles/sweep_arenas.sh -- env-step throughput vs arenas per GPU (run under gpurun)
for n in 8192 16384 32768 65536 131072 262144; do
  timeout 200 python bench.py --arenas $n --steps 100 --warmup 10 --no-cpu-baseline --no-rollout 2>&1 | tail -1 > /tmp/l.json
  python - "$n" <<'PY'
import json, sys
d = json.load(open('/tmp/l.json'))
print(sys.argv[1], "arenas:", round(d["value"] / 1e6, 1), "M env-steps/s,", round(d["ms_per_step"] * 1000, 1), "us/step, roofline frac",
      round(d["roofline"]["frac"], 4), ", e2e", round(d["e2e"]["value"] / 1e6, 1), "M/s")
PY
done
