#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "fixed_operand" 2>&1 | tail -8
timeout 300 python bench.py --workload cl8_fixed_gp_f64 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/final_cl8_fixed_gp_f64.json 2> gpurun_out/final_cl8_f64.err; tail -2 gpurun_out/final_cl8_f64.err
python -c "import json; d=json.load(open('gpurun_out/final_cl8_fixed_gp_f64.json')); r=d['roofline']; print(d['ms_per_step'], d['value']/1e9, r['achieved'], r['frac'], r['hbm_gbs'], d['parity'], d['clocks'])"
LCC_DISABLE_TENSOR_PATH=1 timeout 300 python bench.py --workload cl8_fixed_gp_f64 --log2-rows 20 --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('table kernel, 2^20 rows: ms', d['ms_per_step'], 'TFLOP/s', d['roofline']['achieved'])"
