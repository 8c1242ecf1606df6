#!/bin/bash
# Parity of the TMA-staged SoA kernels + A/B bandwidth (LCC_SOA_TMA=0 = direct-load kernels).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -8
for wl in pga_sandwich cga_gp cga_gp_outer r3_gp; do
  for tma in 0 1; do
    LCC_SOA_TMA=$tma timeout 300 python bench.py --workload $wl --no-e2e --no-cpu --steps 10 > gpurun_out/soa_${wl}_tma$tma.json 2> gpurun_out/soa_${wl}_tma$tma.err
    echo "$wl tma=$tma rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/soa_${wl}_tma$tma.json')); print(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['parity']['ok'], d['clocks'])" 2>&1 | tail -1)"
  done
done
