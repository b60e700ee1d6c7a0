#!/bin/bash
# profiles/run_r1n.sh -- full pass with the fused policy-forward kernel in the sampler: tests, bench, launch list, ncu of both kernels
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r1n.log
echo "== bench (default)"
timeout 900 python bench.py 2> gpurun_out/bench_r1n.err | tail -1 > gpurun_out/bench_r1n.json; cat gpurun_out/bench_r1n.json
echo "== policy forward probe"
timeout 200 python profiles/policy_forward_probe.py 2>&1 | tee gpurun_out/r1n_policy_forward.txt
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1n.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-hier > gpurun_out/ncu_list.log 2>&1
echo "== ncu full (policy forward, 3xTF32)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_forward -s 6 -c 1 -f -o gpurun_out/prof_policy_r1n python profiles/policy_forward_probe.py > gpurun_out/ncu_policy.log 2>&1
ls -la gpurun_out | tail -8
