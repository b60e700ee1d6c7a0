#!/bin/bash
# profiles/run_r3h.sh -- end-of-round ncu evidence: launch list of the (grouped) rollout leg, --set full capture of the policy forward
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_r3h_rollout.csv \
  python bench.py --leg rollout --steps 40 --warmup 3 --no-cpu-baseline --no-hier --no-l5 --no-ppo > gpurun_out/ncu_r3h_list.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:policy_forward_m128 -s 6 -c 2 -o gpurun_out/prof_policy_r3h -f \
  python profiles/policy_forward_probe.py 8192 > gpurun_out/ncu_r3h_full.log 2>&1
tail -2 gpurun_out/ncu_r3h_full.log
