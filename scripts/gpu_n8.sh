#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"; free -g | head -2; nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 tests/tools/multi_gpu_check.py > gpurun_out/multi_gpu_check_n$NG.log 2>&1; echo "multi rc=$?"; grep "rank" gpurun_out/multi_gpu_check_n$NG.log | sort | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err; echo "bench rc=$?"; cat gpurun_out/bench_n$NG.json | cut -c1-1800; tail -3 gpurun_out/bench_n$NG.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $NG --workload sta_fused_grade_inner_sum --no-e2e --no-cpu > gpurun_out/bench_sta_fused_n$NG.json 2> gpurun_out/bench_sta_fused_n$NG.err; echo "sta rc=$?"; cat gpurun_out/bench_sta_fused_n$NG.json | cut -c1-600
timeout 300 python bench.py --impl reference --gpus $NG --steps 3 --warmup 1 | cut -c1-400
