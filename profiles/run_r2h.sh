#!/bin/bash
# profiles/run_r2h.sh -- step kernel: arenas per CTA (one CTA per SM at 8 192 arenas = 56) x rolled row-copy loops
mkdir -p gpurun_out
for lib in "" "$PWD/build/lib_v4rolled.so" "$PWD/build/lib_v4a48r.so" "$PWD/build/lib_v4a56r.so" "$PWD/build/lib_v4a64r.so"; do
  for n in 8192 32768 131072; do
    HH_LIB_PATH=$lib timeout 200 python bench.py --arenas $n --steps 200 --warmup 20 --no-cpu-baseline --no-rollout --no-hier --no-l5 --no-ppo 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.load(sys.stdin); print('${lib:-default}', d['config']['arenas_per_gpu'], round(d['value']/1e6,1), 'M env-steps/s', round(d['ms_per_step']*1000,2), 'us', 'median', round(d['step_time']['per_rank_us'][0]['median'],2), 'e2e', round(d['e2e']['value']/1e6,1))"
  done
done | tee gpurun_out/r2h_arenas_per_cta.txt
