#!/bin/bash
# Persistent three-stage DMMA kernel A/B, the host entry point for compact points (test + e2e line).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fp64_tensor_pipe or host_pga_points or pga_points_fused or host_buffer" > gpurun_out/pytest_r10.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r10.log
for p in 0 1; do
  LCC_DMMA_PERSISTENT=$p timeout 300 python bench.py --workload cl8_fixed_gp_f64 --no-e2e --no-cpu --steps 10 > gpurun_out/dmma_persistent_$p.json 2> gpurun_out/dmma_persistent_$p.err
  python -c "import json; d=json.load(open('gpurun_out/dmma_persistent_$p.json')); r=d['roofline']; print('persistent $p', round(d['ms_per_step'],3),'ms', round(r['achieved'],2), r['unit'], round(r['frac'],3), d['parity']['ok'], d['clocks'])" 2>&1 | tail -1
done
timeout 600 python bench.py --workload pga_points_compact --steps 20 > gpurun_out/final_pga_points_compact.json 2> gpurun_out/final_pga_points_compact.err; echo "points rc=$?"; python -c "import json; d=json.load(open('gpurun_out/final_pga_points_compact.json')); print(d['value']/1e9, d['roofline']['achieved'], d['e2e'], d['cpu_baseline'])"; tail -3 gpurun_out/final_pga_points_compact.err
