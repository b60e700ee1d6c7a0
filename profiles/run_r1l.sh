#!/bin/bash
# profiles/run_r1l.sh -- v4.1 final pass of this session: tests, bench, ncu launch list + full capture, stage clocks
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r1l.log
echo "== bench (default)"
timeout 600 python bench.py 2> gpurun_out/bench_r1l.err | tail -1 > gpurun_out/bench_r1l.json; cat gpurun_out/bench_r1l.json
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_r1l_reference.json; cat gpurun_out/bench_r1l_reference.json
echo "== stage clocks"
HH_LIB_PATH=$PWD/build/lib_v4prof.so timeout 200 python profiles/stage_clocks.py 8192 2>&1 | tee gpurun_out/r1l_stage_clocks_8192.txt
echo "== sweep"
for n in 8192 32768 131072; do
  timeout 200 python bench.py --arenas $n --steps 100 --warmup 10 --no-cpu-baseline --no-rollout --no-hier 2>/dev/null | tail -1 > /tmp/l.json
  python - "v4.1" "$n" <<'PY' | tee -a gpurun_out/r1l_sweep.txt
import json, sys
d = json.load(open('/tmp/l.json'))
print(sys.argv[1], sys.argv[2], "arenas:", round(d["value"] / 1e6, 1), "M env-steps/s,", round(d["ms_per_step"] * 1000, 1), "us/step, b2b",
      round(d["back_to_back"]["value"] / 1e6, 1), "M, e2e", round(d["e2e"]["value"] / 1e6, 1), "M, other host mode", round(d["e2e"]["other_host_mode"]["value"] / 1e6, 1), "M")
PY
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1l.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel_v4 -s 5 -c 2 -f -o gpurun_out/prof_step_r1l python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-rollout --no-hier > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -12
