#!/bin/bash
# K3 with two chunks per thread, 64-thread normalized (SoA), and the blade-row pitch experiment (is a power-of-two
# distance between the blade rows of an SoA batch slower than a padded one?).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_batch_api.py -x -q -k "unary or normalized or norm or dual or reverse or grade or scale" > gpurun_out/pytest_k3.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_k3.log
timeout 600 python scripts/bench_aux_ops.py > gpurun_out/aux_ops_r7.jsonl 2> gpurun_out/aux_ops_r7.err; echo "aux rc=$?"
grep -E "reverse|grade1|scale|dual|normalized" gpurun_out/aux_ops_r7.jsonl | cut -c1-200
for wl in pga_sandwich cga_gp; do
  for pad in 0 64 4160 65600; do
    timeout 300 python bench.py --workload $wl --ld-pad $pad --no-e2e --no-cpu --steps 20 > gpurun_out/ldpad_${wl}_$pad.json 2> gpurun_out/ldpad_${wl}_$pad.err
    python -c "import json; d=json.load(open('gpurun_out/ldpad_${wl}_$pad.json')); r=d['roofline']; print('$wl pad $pad', round(d['ms_per_step'],4),'ms', round(r['achieved'],1), r['unit'], d['parity']['ok'])" 2>&1 | tail -1
  done
done
