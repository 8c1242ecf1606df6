#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
timeout 300 python scripts/bench_torch_interop.py > gpurun_out/torch_interop.jsonl 2> gpurun_out/torch_interop.err; echo "interop rc=$?"; cat gpurun_out/torch_interop.jsonl; tail -3 gpurun_out/torch_interop.err
CL8="python bench.py --workload cl8_fixed_gp --steps 10 --warmup 3 --no-e2e"
timeout 600 $CL8 > gpurun_out/bench_cl8.json 2> gpurun_out/bench_cl8.err; echo "cl8 rc=$?"; cat gpurun_out/bench_cl8.json; tail -3 gpurun_out/bench_cl8.err
CL8P="python bench.py --workload cl8_fixed_gp --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity"
$CL8P > gpurun_out/plain_cl8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fixed_product -s 3 -c 1 -o gpurun_out/prof_cl8_fixed $CL8P > gpurun_out/ncu_cl8.log 2>&1
echo "ncu rc=$?"
CGA="python bench.py --workload cga_gp --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity"
$CGA > gpurun_out/plain_cga.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:product_kernel -s 3 -c 1 -o gpurun_out/prof_cga_gp_v2 $CGA > gpurun_out/ncu_cga.log 2>&1
echo "ncu cga rc=$?"
PGA="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity"
$PGA > gpurun_out/plain_pga.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sandwich_kernel -s 3 -c 1 -o gpurun_out/prof_pga_v2 $PGA > gpurun_out/ncu_pga.log 2>&1
echo "ncu pga rc=$?"
