#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
{ nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|model name\|^cpu(s)"; numactl -H 2>/dev/null | head -20; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; } > gpurun_out/topo_n$NG.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29551 scripts/e2e_numa_probe.py > gpurun_out/numa_probe_n$NG.log 2>&1; echo "probe rc=$?"
grep "^rank" gpurun_out/numa_probe_n$NG.log | sort | head -20
head -30 gpurun_out/topo_n$NG.txt
