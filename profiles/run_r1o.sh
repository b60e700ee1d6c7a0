#!/bin/bash
# profiles/run_r1o.sh -- v4.2 (cold paths out of line, reset draws by 8 threads, state store merged into the row stage)
mkdir -p gpurun_out
echo "== pytest -m gpu (parity)"
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r1o.log
echo "== stage clocks"
HH_LIB_PATH=$PWD/build/lib_v4prof.so timeout 200 python profiles/stage_clocks.py 8192 2>&1 | tee gpurun_out/r1o_stage_clocks_8192.txt
echo "== sweep"
for n in 8192 32768 131072; do
  timeout 200 python bench.py --arenas $n --steps 200 --warmup 20 --no-cpu-baseline --no-rollout --no-hier --no-l5 2>/dev/null | tail -1 > /tmp/l.json
  python - "v4.2" "$n" <<'PY' | tee -a gpurun_out/r1o_sweep.txt
import json, sys
d = json.load(open('/tmp/l.json'))
print(sys.argv[1], sys.argv[2], "arenas:", round(d["value"] / 1e6, 1), "M env-steps/s,", round(d["ms_per_step"] * 1000, 1), "us/step, b2b",
      round(d["back_to_back"]["value"] / 1e6, 1), "M, e2e", round(d["e2e"]["value"] / 1e6, 1), "M, other host mode", round(d["e2e"]["other_host_mode"]["value"] / 1e6, 1), "M")
PY
done
