#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -60 gpurun_out/pytest.log
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity"
$BENCH > gpurun_out/plain_pga.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pga.csv $BENCH > gpurun_out/ncu_pga.log 2>&1
echo "ncu launches rc=$?"
$BENCH > gpurun_out/plain_pga2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sandwich_kernel -s 3 -c 2 -o gpurun_out/prof_pga_sandwich $BENCH > gpurun_out/ncu_pga_full.log 2>&1
echo "ncu full pga rc=$?"
CGA="python bench.py --workload cga_gp --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity"
$CGA > gpurun_out/plain_cga.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:product_kernel -s 3 -c 2 -o gpurun_out/prof_cga_gp $CGA > gpurun_out/ncu_cga_full.log 2>&1
echo "ncu full cga rc=$?"
ls -la gpurun_out
