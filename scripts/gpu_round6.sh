#!/bin/bash
# After the second batch of changes: full GPU suite, aux-op bandwidths, ncu captures of the compact-points kernel
# and the TMA transpose, closing bench lines.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
timeout 600 python scripts/bench_aux_ops.py > gpurun_out/aux_ops_r6.jsonl 2> gpurun_out/aux_ops_r6.err; echo "aux rc=$?"
grep -E "sum|norm|convert" gpurun_out/aux_ops_r6.jsonl | cut -c1-200
timeout 300 python bench.py --workload pga_points_compact --no-e2e --steps 20 > gpurun_out/final_pga_points_compact.json 2> gpurun_out/final_pga_points_compact.err; echo "points rc=$?"; cut -c1-700 gpurun_out/final_pga_points_compact.json
PTS="python bench.py --workload pga_points_compact --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pga_points_sandwich -s 3 -c 1 -f -o gpurun_out/prof_pga_points $PTS > gpurun_out/ncu_points.log 2>&1; echo "ncu points rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:transpose_tma -s 3 -c 2 -f -o gpurun_out/prof_transpose_tma python scripts/bench_aux_ops.py convert > gpurun_out/ncu_transpose.log 2>&1; echo "ncu transpose rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:norm_aos_lanes -s 3 -c 1 -f -o gpurun_out/prof_norm_aos python scripts/bench_aux_ops.py norm > gpurun_out/ncu_norm.log 2>&1; echo "ncu norm rc=$?"
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/final_bench_n1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference_n1.json 2>/dev/null; echo "ref rc=$?"
ls -la gpurun_out | tail -20
