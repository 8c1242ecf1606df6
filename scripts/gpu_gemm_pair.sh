#!/bin/bash
# K6 on CTA pairs: parity first (short timeout: a barrier mistake hangs), then A/B against the one-CTA kernel.
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -q -x -k "fixed_operand" 2>&1 | tail -12
echo "pytest rc=${PIPESTATUS[0]}"
for pair in 0 1; do
  LCC_GEMM_PAIR=$pair timeout 240 python bench.py --workload cl8_fixed_gp --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/cl8_pair$pair.json 2> gpurun_out/cl8_pair$pair.err
  echo "pair=$pair rc=$? $(python -c "import json; d=json.load(open('gpurun_out/cl8_pair$pair.json')); r=d['roofline']; print(d['ms_per_step'], r['achieved'], r['frac'], r['hbm_gbs'], d['parity'], d['clocks'])" 2>&1 | tail -1)"
done
for pairs in 60 66 70 72; do
  LCC_GEMM_PAIRS=$pairs timeout 240 python bench.py --workload cl8_fixed_gp --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/cl8_pairs$pairs.json 2> gpurun_out/cl8_pairs$pairs.err
  echo "pairs=$pairs rc=$? $(python -c "import json; d=json.load(open('gpurun_out/cl8_pairs$pairs.json')); r=d['roofline']; print(d['ms_per_step'], r['achieved'], r['frac'], d['clocks'])" 2>&1 | tail -1)"
done
