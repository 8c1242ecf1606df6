#!/bin/bash
# SoA TMA tile-size experiment: quick parity + bandwidth of the four SoA workloads with the library as built.
mkdir -p gpurun_out
TAG=${1:-exp}
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "soa_batches_through_tma or sandwich" 2>&1 | tail -3
for wl in pga_sandwich cga_gp cga_gp_outer r3_gp; do
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu --steps 10 > gpurun_out/soatile_${TAG}_${wl}.json 2> gpurun_out/soatile_${TAG}_${wl}.err
    echo "$TAG $wl rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/soatile_${TAG}_${wl}.json')); print(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['parity']['ok'], d['clocks'])" 2>&1 | tail -1)"
done
