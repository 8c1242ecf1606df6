#!/bin/bash
# Round-1 closing numbers at N=1: the default bench line, every other workload, both layouts, and the reference arm.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/final_bench_n1.json
for wl in cga_gp cga_gp_outer r3_gp sta_grade_inner_sum sta_fused_grade_inner_sum cl8_fixed_gp; do
  timeout 400 python bench.py --workload $wl --no-e2e --no-cpu --steps 10 > gpurun_out/final_$wl.json 2> gpurun_out/final_$wl.err
  echo "$wl rc=$? $(python -c "import json; d=json.load(open('gpurun_out/final_$wl.json')); r=d['roofline']; print(round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],3),'ms', round(r['achieved'],1), r['unit'], round(r['frac'],3), d['parity']['ok'], d['clocks']['reasons'])" 2>&1 | tail -1)"
done
for wl in pga_sandwich cga_gp cga_gp_outer r3_gp; do
  timeout 400 python bench.py --workload $wl --layout aos --no-e2e --no-cpu --steps 10 > gpurun_out/final_aos_$wl.json 2> gpurun_out/final_aos_$wl.err
  echo "aos $wl rc=$? $(python -c "import json; d=json.load(open('gpurun_out/final_aos_$wl.json')); r=d['roofline']; print(round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],3),'ms', round(r['achieved'],1), r['unit'], round(r['frac'],3), d['parity']['ok'])" 2>&1 | tail -1)"
done
timeout 400 python bench.py --workload r3_gp --log2-rows 26 --no-e2e --no-cpu --steps 10 > gpurun_out/final_r3_gp_2p26.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/final_r3_gp_2p26.json')); print('r3 2^26', d['roofline']['achieved'], d['roofline']['frac'])"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference_n1.json 2>/dev/null; cut -c1-500 gpurun_out/final_reference_n1.json
timeout 300 python scripts/bench_torch_interop.py > gpurun_out/final_torch_interop.jsonl 2> gpurun_out/final_torch_interop.err; cat gpurun_out/final_torch_interop.jsonl | cut -c1-300
