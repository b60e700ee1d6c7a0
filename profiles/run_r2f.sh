#!/bin/bash
# profiles/run_r2f.sh -- full GPU suite + default bench with the tcgen05 forward as the default policy path
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r2f.log
timeout 120 python profiles/policy_forward_probe.py 8192 2>&1 | head -4 | tee gpurun_out/r2f_policy_forward.txt
timeout 900 python bench.py > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; tail -c 3000 gpurun_out/bench_r2f.json
