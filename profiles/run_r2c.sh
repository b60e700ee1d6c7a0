#!/bin/bash
# profiles/run_r2c.sh -- first run of the tcgen05 policy forward (csrc/hh_policy_tc.cu): layout test, parity, timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampler.py -m gpu -x -q -k "pack_image or fused_policy_forward or actor_chains" 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r2c.log
timeout 300 python profiles/policy_forward_probe.py 8192 2>&1 | tail -8 | tee gpurun_out/r2c_policy_forward.txt
