#!/bin/bash
# Final N=1 lines on the final tree (stdout must carry exactly one JSON line), launch list of the bench command, and an
# ncu look at the AoS batch-sum kernel.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/final_bench_n1.json)"; cut -c1-200 gpurun_out/final_bench_n1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference_n1.json 2>/dev/null; echo "ref rc=$? lines=$(wc -l < gpurun_out/final_reference_n1.json)"
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_final.csv $BENCH > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sum_partial_aos_vec -s 3 -c 1 -f -o gpurun_out/prof_sum_aos python scripts/bench_aux_ops.py sum > gpurun_out/ncu_sum.log 2>&1; echo "ncu sum rc=$?"
