#!/bin/bash
# profiles/run_r1j.sh -- stage clocks of the v4 kernel, 64-arena CTAs, exact short fmod (run under gpurun)
mkdir -p gpurun_out
one() {  # label, env..., arenas
  local label="$1"; shift
  local n="$1"; shift
  env "$@" timeout 200 python bench.py --arenas $n --steps 100 --warmup 10 --no-cpu-baseline --no-rollout --no-hier 2>/dev/null | tail -1 > /tmp/l.json
  python - "$label" "$n" <<'PY' | tee -a gpurun_out/r1j_sweep.txt
import json, sys
d = json.load(open('/tmp/l.json'))
print(sys.argv[1], sys.argv[2], "arenas:", round(d["value"] / 1e6, 1), "M env-steps/s,", round(d["ms_per_step"] * 1000, 1), "us/step, b2b",
      round(d["back_to_back"]["value"] / 1e6, 1), "M, e2e", round(d["e2e"]["value"] / 1e6, 1), "M, other host mode", round(d["e2e"]["other_host_mode"]["value"] / 1e6, 1), "M")
PY
}
for n in 8192 32768 131072; do one "v4 32-arena CTAs + short fmod" $n HH_DUMMY=1; done
for n in 8192 32768 131072; do one "v4 64-arena CTAs + short fmod" $n HH_LIB_PATH=$PWD/build/lib_v4a64.so; done
echo "== parity of the 64-arena build"
HH_LIB_PATH=$PWD/build/lib_v4a64.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "oracle_many or full_size or golden or zero_copy" 2>&1 | tail -2
echo "== stage clocks"
HH_LIB_PATH=$PWD/build/lib_v4prof.so timeout 200 python profiles/stage_clocks.py 8192 2>&1 | tee gpurun_out/r1j_stage_clocks_8192.txt
HH_LIB_PATH=$PWD/build/lib_v4prof.so timeout 200 python profiles/stage_clocks.py 131072 2>&1 | tee gpurun_out/r1j_stage_clocks_131072.txt
