#!/bin/bash
# profiles/run_r2a.sh -- FIRST GPU call of round 2 (prepared at the end of round 1, when the GPU budget was spent):
#   1. the GPU tests that were added after the last GPU run of round 1 (evaluation statistics, trace recorder);
#   2. the tcgen05 / TMEM probe (profiles/tcgen05_probe.cu; binary prebuilt by profiles/build_variants.sh into build/):
#      descriptor layout, TMEM mapping for M = 64 / 128, operand narrowing, 3xTF32 accuracy, MMA issue rate;
#   3. the instruction-fetch analysis of a fresh step-kernel capture (profiles/stall_by_address.py, code_by_stage.py).
mkdir -p gpurun_out
echo "== new GPU tests"
timeout 600 python -m pytest tests/test_gpu_hier.py tests/test_gpu_zz_evaluation.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r2a.log
echo "== tcgen05 probe"
timeout 120 build/tcgen05_probe 2>&1 | tee gpurun_out/r2a_tcgen05_probe.txt
if grep -q "rows not found [1-9]\|did not complete" gpurun_out/r2a_tcgen05_probe.txt; then
  echo "== tcgen05 probe, descriptor offsets swapped"
  timeout 120 build/tcgen05_probe swap 2>&1 | tee gpurun_out/r2a_tcgen05_probe_swap.txt
fi
echo "== step kernel: default vs the instruction-footprint variants (profiles/build_variants.sh)"
for lib in "" "$PWD/build/lib_v4warm.so" "$PWD/build/lib_v4rolled.so" "$PWD/build/lib_v4rolled_warm.so"; do
  for n in 8192 131072; do
    HH_LIB_PATH=$lib timeout 200 python bench.py --arenas $n --steps 200 --warmup 20 --no-cpu-baseline --no-rollout --no-hier --no-l5 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.load(sys.stdin); print('${lib:-default}', d['config']['arenas_per_gpu'], round(d['value']/1e6,1), 'M env-steps/s', round(d['ms_per_step']*1000,2), 'us')"
  done
done | tee gpurun_out/r2a_warm_variant.txt
echo "== step kernel capture"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel_v4 -s 5 -c 2 -o gpurun_out/prof_step_r2a -f \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-rollout --no-hier --no-l5 > gpurun_out/ncu_r2a.log 2>&1
python profiles/stall_by_address.py gpurun_out/prof_step_r2a.ncu-rep 2>&1 | tee gpurun_out/r2a_step_kernel_no_instruction.txt
