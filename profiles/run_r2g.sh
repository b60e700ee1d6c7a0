#!/bin/bash
# profiles/run_r2g.sh -- full GPU suite (incl. the L4/L5 golden replays, escape-mode sampler, compute_actions, learning test)
# + the default bench and its rollout leg
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r2g.log
timeout 900 python bench.py > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; tail -c 1500 gpurun_out/bench_r2g.err
timeout 600 python bench.py --leg rollout --steps 100 --no-hier --no-l5 --no-ppo > gpurun_out/bench_r2g_rollout.json 2>> gpurun_out/bench_r2g.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_r2g_reference.json 2>> gpurun_out/bench_r2g.err
