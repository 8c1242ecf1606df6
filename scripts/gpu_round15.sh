#!/bin/bash
# SoA `normalized`: register kernel with 16-byte accesses (0), with 8-byte accesses (1), TMA tiles (2)
mkdir -p gpurun_out
for m in 0 1 2; do
  LCC_NORMALIZED_MODE=$m timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "normalized_soa" > gpurun_out/pytest_norm_$m.log 2>&1; echo "mode $m pytest rc=$?"
  LCC_NORMALIZED_MODE=$m AUX_ALL_SIGS=1 timeout 300 python scripts/bench_aux_ops.py normalized 2>/dev/null | grep '"soa"' > gpurun_out/aux_normalized_$m.jsonl; sed "s/^/mode $m /" gpurun_out/aux_normalized_$m.jsonl | cut -c1-170
done
