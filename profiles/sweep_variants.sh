#!/bin/bash
# profiles/sweep_variants.sh -- CTA size / phase-barrier variants of the step kernel (run under gpurun)
for v in t32_s0 t128_s0 t128_s1 t256_s1; do
  for n in 8192 32768; do
    HH_LIB_PATH=$PWD/build/lib_$v.so timeout 200 python bench.py --arenas $n --steps 100 --warmup 10 --no-cpu-baseline --no-rollout 2>&1 | tail -1 > /tmp/l.json
    python - "$v" "$n" <<'PY'
import json, sys
d = json.load(open('/tmp/l.json'))
print(sys.argv[1], sys.argv[2], "arenas:", round(d["value"] / 1e6, 1), "M env-steps/s,", round(d["ms_per_step"] * 1000, 1), "us/step")
PY
  done
done
