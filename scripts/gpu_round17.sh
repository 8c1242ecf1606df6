#!/bin/bash
# Closing check on the final tree: smoke, the whole GPU suite, the aux-op table, the default bench line.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 600 python scripts/bench_aux_ops.py > gpurun_out/aux_ops_final.jsonl 2>/dev/null; echo "aux rc=$? lines=$(wc -l < gpurun_out/aux_ops_final.jsonl)"
timeout 900 python bench.py --no-cpu > gpurun_out/bench_check.json 2>/dev/null; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench_check.json)"; cut -c1-160 gpurun_out/bench_check.json
