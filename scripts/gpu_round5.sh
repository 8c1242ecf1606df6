#!/bin/bash
# Re-entry check of the rebuilt tree + A/B of the latency fixes (vectorised SoA sum, 4-deep AoS norm), the TMA transpose
# and the block shapes of the compact-points kernel.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
for shape in 0 1 2; do
  LCC_POINTS_SHAPE=$shape timeout 300 python -m pytest tests/test_gpu_parity.py -q -k pga_points_fused > gpurun_out/pytest_points_$shape.log 2>&1; echo "points shape $shape pytest rc=$?"
  LCC_POINTS_SHAPE=$shape timeout 300 python bench.py --workload pga_points_compact --no-e2e --no-cpu --steps 20 > gpurun_out/points_shape_$shape.json 2> gpurun_out/points_shape_$shape.err
  python -c "import json; d=json.load(open('gpurun_out/points_shape_$shape.json')); r=d['roofline']; print('shape $shape', round(d['value']/1e9,2),'G/s', round(d['ms_per_step'],3),'ms', round(r['achieved'],1), r['unit'], d['parity']['ok'])" 2>&1 | tail -1
done
timeout 600 python scripts/bench_aux_ops.py > gpurun_out/aux_ops_r5.jsonl 2> gpurun_out/aux_ops_r5.err; echo "aux rc=$?"
LCC_TMA_TRANSPOSE=0 timeout 600 python scripts/bench_aux_ops.py 2>/dev/null | grep convert > gpurun_out/aux_ops_r5_scalar_transpose.jsonl
grep -E "sum|norm\"|convert" gpurun_out/aux_ops_r5.jsonl | cut -c1-200
echo "--- scalar transpose"; cat gpurun_out/aux_ops_r5_scalar_transpose.jsonl | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench_n1_r5.json 2> gpurun_out/bench_n1_r5.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_n1_r5.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/reference_n1_r5.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 gpurun_out/reference_n1_r5.json
