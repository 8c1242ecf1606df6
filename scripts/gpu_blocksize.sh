#!/bin/bash
# block-size sweep for the register-heavy kernels (LCC_BLOCK_THREADS override)
mkdir -p gpurun_out
for wl in cga_gp cga_gp_outer pga_sandwich; do
  for bt in 64 128 256; do
    LCC_BLOCK_THREADS=$bt timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl block=$bt', round(d['ms_per_step'],3),'ms', round(d['roofline']['achieved'],1),'GB/s', round(d['roofline']['frac'],4), d['clocks'])"
  done
done | tee gpurun_out/blocksize.log
