#!/bin/bash
# profiles/run_r1k.sh -- v4.1 (single arccos site, candidate-compacted cannon geometry, trimmed direct_short, 3-thread rows)
mkdir -p gpurun_out
one() {  # label, arenas, env...
  local label="$1"; shift
  local n="$1"; shift
  env "$@" timeout 200 python bench.py --arenas $n --steps 100 --warmup 10 --no-cpu-baseline --no-rollout --no-hier 2>/dev/null | tail -1 > /tmp/l.json
  python - "$label" "$n" <<'PY' | tee -a gpurun_out/r1k_sweep.txt
import json, sys
d = json.load(open('/tmp/l.json'))
print(sys.argv[1], sys.argv[2], "arenas:", round(d["value"] / 1e6, 1), "M env-steps/s,", round(d["ms_per_step"] * 1000, 1), "us/step, b2b",
      round(d["back_to_back"]["value"] / 1e6, 1), "M, e2e", round(d["e2e"]["value"] / 1e6, 1), "M, other host mode", round(d["e2e"]["other_host_mode"]["value"] / 1e6, 1), "M")
PY
}
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r1k.log
for n in 8192 32768 131072; do one "v4.1 (2 CTAs/SM, 128 regs)" $n HH_DUMMY=1; done
for n in 8192 32768 131072; do one "v4.1 (3 CTAs/SM, 80 regs)" $n HH_LIB_PATH=$PWD/build/lib_v4occ3.so; done
echo "== stage clocks"
HH_LIB_PATH=$PWD/build/lib_v4prof.so timeout 200 python profiles/stage_clocks.py 8192 2>&1 | tee gpurun_out/r1k_stage_clocks_8192.txt
