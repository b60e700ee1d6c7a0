#!/bin/bash
# profiles/run_r2zc.sh -- PPOLearner with CUDA-graph minibatches: tests, timing, launch list of one update
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampler.py -m gpu -q -k "graph_replayed or learner_improves or escape_mode" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r2zc.log
timeout 300 python profiles/ppo_learner_probe.py 2>&1 | grep use_cuda_graph | tee gpurun_out/r2zc_ppo_learner.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20000 -c 1500 --csv --log-file gpurun_out/launches_r2zc_learner.csv \
  python profiles/ppo_learner_probe.py 2048 > gpurun_out/ncu_r2zc_list.log 2>&1
tail -2 gpurun_out/ncu_r2zc_list.log
