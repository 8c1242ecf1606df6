#!/bin/bash
# AoS norm / normalized on TMA tiles vs the lane-group kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "aos_views or normalized" > gpurun_out/pytest_r16.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r16.log
for m in 0 1; do
  LCC_NORM_AOS_TMA=$m AUX_ALL_SIGS=1 timeout 300 python scripts/bench_aux_ops.py norm 2>/dev/null | grep '"aos"' > gpurun_out/aux_norm_aos_tma_$m.jsonl; sed "s/^/tma $m /" gpurun_out/aux_norm_aos_tma_$m.jsonl | cut -c1-170
done
