#!/bin/bash
# Randomised strict-mode parity sweep on the final tree (now also layout conversion and batch sums)
mkdir -p gpurun_out
for seed in 13 14; do
  FUZZ_SEED=$seed timeout 75 python tests/tools/fuzz_parity.py 2000 > gpurun_out/fuzz_$seed.log 2>&1; echo "seed $seed rc=$?"; tail -2 gpurun_out/fuzz_$seed.log | cut -c1-300
done
