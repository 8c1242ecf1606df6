#!/bin/bash
# Two GPUs of one box: sharded headline bench (with the PCIe end-to-end leg), fused STA workload with the NCCL batch sum,
# and the shard-by-shard parity check.
mkdir -p gpurun_out
NG=2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29541 tests/tools/multi_gpu_check.py > gpurun_out/multi_gpu_check_n$NG.log 2>&1; echo "multi rc=$?"; grep "rank" gpurun_out/multi_gpu_check_n$NG.log | sort | head -4
timeout 600 $RUN --master-port 29542 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/bench_pga_n$NG.json 2> gpurun_out/bench_pga_n$NG.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_pga_n$NG.json
timeout 300 python bench.py --workload pga_points_compact --steps 20 --no-cpu > gpurun_out/final_pga_points_compact.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/final_pga_points_compact.json')); print('points e2e', d['e2e']['value']/1e9)"
timeout 600 $RUN --master-port 29543 bench.py --gpus $NG --steps 20 --warmup 3 --workload sta_fused_grade_inner_sum --no-e2e > gpurun_out/bench_sta_fused_n$NG.json 2> gpurun_out/bench_sta_fused_n$NG.err; echo "sta rc=$?"; cut -c1-200 gpurun_out/bench_sta_fused_n$NG.json
