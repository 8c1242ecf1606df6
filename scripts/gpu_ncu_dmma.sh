#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload cl8_fixed_gp_f64 --log2-rows 22 --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity"
$CMD > gpurun_out/plain_dmma.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dmma_kernel -s 3 -c 1 -f -o gpurun_out/prof_cl8_dmma $CMD > gpurun_out/ncu_dmma.log 2>&1
echo "ncu rc=$?"
