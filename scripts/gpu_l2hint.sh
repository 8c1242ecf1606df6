#!/bin/bash
# A/B of the L2 evict-first hint on the TMA copies (LCC_TMA_L2_HINT=1) for the headline and CGA workloads, both layouts.
mkdir -p gpurun_out
for hint in 0 1; do
  for args in "--workload pga_sandwich" "--workload pga_sandwich --layout aos" "--workload cga_gp" "--workload cga_gp --layout aos" "--workload cga_gp_outer"; do
    LCC_TMA_L2_HINT=$hint timeout 300 python bench.py $args --no-e2e --no-cpu --steps 10 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('hint=$hint', '$args', round(d['ms_per_step'],4), round(d['roofline']['achieved'],1), d['parity']['ok'], d['clocks']['reasons'])"
  done
done
