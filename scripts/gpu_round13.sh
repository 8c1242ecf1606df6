#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "sum or aos_views or sharded" > gpurun_out/pytest_r13.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_r13.log
AUX_ALL_SIGS=1 timeout 300 python scripts/bench_aux_ops.py sum > gpurun_out/aux_sum_r13.jsonl 2>/dev/null; cat gpurun_out/aux_sum_r13.jsonl | cut -c1-160
