#!/bin/bash
# profiles/run_r2e.sh -- tcgen05 policy forward with cluster multicast of the weight stream: parity, timing per cluster size, stamps
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampler.py -m gpu -x -q -k "pack_image or fused_policy_forward or actor_chains" 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r2e.log
timeout 300 python profiles/policy_forward_probe.py 8192 2>&1 | tail -10 | tee gpurun_out/r2e_policy_forward.txt
for cs in 2; do timeout 120 python profiles/tc_profile.py 8192 $cs 2>&1 | tail -34; done | tee gpurun_out/r2e_tc_profile.txt
