#!/bin/bash
# Full GPU suite including the full-size property tests, the fused STA line with the finer default grid, small-batch latency.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/pytest.log
timeout 300 python bench.py --workload sta_fused_grade_inner_sum --no-e2e --no-cpu --steps 20 > gpurun_out/final_sta_fused_grade_inner_sum.json 2>/dev/null; cut -c1-250 gpurun_out/final_sta_fused_grade_inner_sum.json; python -c "import json; d=json.load(open('gpurun_out/final_sta_fused_grade_inner_sum.json')); print(d['ms_per_step'], d['roofline'])"
timeout 300 python scripts/bench_small_batches.py > gpurun_out/small_batches.jsonl 2> gpurun_out/small_batches.err; echo "small rc=$?"; cat gpurun_out/small_batches.jsonl; tail -3 gpurun_out/small_batches.err
