#!/bin/bash
# profiles/run_r1q.sh -- final pass of round 1: tests, bench (+ reference arm), ncu launch lists, ncu full of the step kernel
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r1q.log
echo "== bench (default)"
timeout 900 python bench.py 2> gpurun_out/bench_r1q.err | tail -1 > gpurun_out/bench_r1q.json; cat gpurun_out/bench_r1q.json | cut -c1-400
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_r1q_reference.json; cut -c1-200 gpurun_out/bench_r1q_reference.json
echo "== ncu launch list (default bench, first 500 launches)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r1q.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-hier --no-l5 > gpurun_out/ncu_list.log 2>&1
echo "== ncu launch list (own kernels of the hierarchical and level-5 legs)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tick_kernel|agents_kernel|begin_kernel|end_kernel|finish_kernel|policy_forward|reset_kernel" -c 400 --csv --log-file gpurun_out/launches_r1q_hier_l5.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-rollout > gpurun_out/ncu_list2.log 2>&1
echo "== ncu full (step kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel_v4 -s 5 -c 2 -f -o gpurun_out/prof_step_r1q python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-rollout --no-hier --no-l5 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
