#!/bin/bash
# profiles/run_r2y.sh -- the 128-row form of the tcgen05 forward (lo halves of the activations in tensor memory)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampler.py -m gpu -x -q -k "pack_image or fused_policy_forward or actor_chains or fused_opponents or variants" 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_r2y.log
timeout 300 python profiles/policy_forward_probe.py 8192 2>&1 | head -3 | tee gpurun_out/r2y_policy_forward.txt
timeout 200 python profiles/tc_profile.py 8192 2>&1 | tee gpurun_out/r2y_tc_profile_m128.txt | grep "==\|seg\|waiting\|epi"
