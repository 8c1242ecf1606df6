#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:normalized_soa_regs -s 3 -c 1 -f -o gpurun_out/prof_normalized_soa python scripts/bench_aux_ops.py normalized > gpurun_out/ncu_normalized.log 2>&1; echo "ncu rc=$?"
