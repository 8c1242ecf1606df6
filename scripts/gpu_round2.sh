#!/bin/bash
# tests + headline bench + 2-GPU scaling + NCCL sum check
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 tests/tools/multi_gpu_check.py > gpurun_out/multi_gpu_check.log 2>&1; echo "multi rc=$?"; grep "rank" gpurun_out/multi_gpu_check.log | head
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err; echo "bench$NG rc=$?"; cat gpurun_out/bench_n$NG.json; tail -3 gpurun_out/bench_n$NG.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $NG --workload sta_grade_inner_sum --no-e2e --no-cpu > gpurun_out/bench_sta_n$NG.json 2> gpurun_out/bench_sta_n$NG.err; echo "sta$NG rc=$?"; cat gpurun_out/bench_sta_n$NG.json
fi
