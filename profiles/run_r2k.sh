#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/policy_forward_probe.py 8192 2>&1 | head -3 | tee gpurun_out/r2k_policy_forward.txt
timeout 120 python profiles/tc_profile.py 8192 2>&1 | tail -64 | tee gpurun_out/r2k_tc_profile.txt
