#!/bin/bash
# profiles/run_r2j.sh -- CTA-pair (cta_group::2) policy forward: layout test, parity, timing, per-role stamps; HH_TC_PAIR=0 for A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampler.py -m gpu -x -q -k "pack_image or fused_policy_forward or actor_chains or fused_opponents" 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r2j.log
timeout 300 python profiles/policy_forward_probe.py 8192 2>&1 | tail -8 | tee gpurun_out/r2j_policy_forward.txt
timeout 120 python profiles/tc_profile.py 8192 2>&1 | tail -34 | tee gpurun_out/r2j_tc_profile.txt
HH_TC_PAIR=0 timeout 300 python profiles/policy_forward_probe.py 8192 2>&1 | head -3 | tee -a gpurun_out/r2j_policy_forward.txt
