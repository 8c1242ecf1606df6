#!/bin/bash
# N=1: host points pipeline with separate stages (test + e2e line), then the whole GPU suite once more on the final tree.
mkdir -p gpurun_out
timeout 600 python bench.py --workload pga_points_compact --steps 20 > gpurun_out/final_pga_points_compact.json 2> gpurun_out/final_pga_points_compact.err; echo "points rc=$?"; python -c "import json; d=json.load(open('gpurun_out/final_pga_points_compact.json')); print(d['value']/1e9, d['roofline']['achieved'], d['e2e']['value']/1e9, d['e2e']['checked_against_oracle'])"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
