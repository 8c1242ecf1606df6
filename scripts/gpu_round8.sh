#!/bin/bash
# Grid / block-size sweep of the fused grade + inner product + batch-sum kernel (BASELINE config 4), and one ncu capture.
mkdir -p gpurun_out
for bt in 64 128; do
  for blocks in 4144 4736 8288 16576 33152; do
    LCC_BLOCK_THREADS=$bt LCC_FUSED_SUM_BLOCKS=$blocks timeout 300 python bench.py --workload sta_fused_grade_inner_sum --no-e2e --no-cpu --steps 20 > gpurun_out/sta_fused_${bt}_$blocks.json 2> gpurun_out/sta_fused_${bt}_$blocks.err
    python -c "import json; d=json.load(open('gpurun_out/sta_fused_${bt}_$blocks.json')); r=d['roofline']; print('threads $bt blocks $blocks', round(d['ms_per_step'],4),'ms', round(r['achieved'],1), r['unit'], d['parity']['ok'])" 2>&1 | tail -1
  done
done
STA="python bench.py --workload sta_fused_grade_inner_sum --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:product_sum_kernel -s 3 -c 1 -f -o gpurun_out/prof_sta_fused $STA > gpurun_out/ncu_sta_fused.log 2>&1; echo "ncu rc=$?"
