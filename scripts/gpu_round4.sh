#!/bin/bash
# Final-state evidence for round 1: launch list of the bench command + one full ncu capture per hot kernel.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity"
cap() {  # name, kernel regex, bench args
  local name=$1 rx=$2; shift 2
  $B "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "$name: plain run failed"; tail -3 gpurun_out/plain_$name.log; return; }
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o gpurun_out/prof_$name $B "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name ncu rc=$?"
}
cap pga_soa_tma sandwich_kernel_soa_tma
cap pga_aos_tma sandwich_kernel_aos_tma --layout aos
cap cga_soa_tma product_kernel_soa_tma --workload cga_gp
cap cga_aos_tma product_kernel_aos_tma --workload cga_gp --layout aos
cap cl8_pair_v2 pair_kernel --workload cl8_fixed_gp
# launch list of the default bench command (cold-cache, serialised: shares matter, not absolutes)
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r01b.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
