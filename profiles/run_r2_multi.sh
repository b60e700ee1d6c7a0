#!/bin/bash
# profiles/run_r2_multi.sh N -- multi-GPU checks under `gpurun --gpus N`: two NCCL ranks end a PPO update with identical
# weights (pytest), then bench.py under torchrun on N GPUs (device-timed value, e2e, rollout, level 5, ppo leg with the
# gradient all-reduce over NVLink)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampler.py -m gpu -x -q -k "nccl" 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_nccl_r2.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
tail -c 600 gpurun_out/bench_r2_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_${N}gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'rollout', d['rollout'].get('fused_tc'), 'l5', d['level5'].get('fused_actors'))
print('ppo', json.dumps(d['ppo'])[:900])
print('step_time', json.dumps(d['step_time']['per_rank_us']))
print('affinity', d['cpu_affinity'])
PY
