#!/bin/bash
# scripts/gpu.sh -- the ONE runner for GPU-box work:  gpurun -- 'bash scripts/gpu.sh <stage> [<stage> ...]'
# Each stage writes its artefacts under gpurun_out/ (merged back by gpurun) and prints a short summary.
# A number printed by a run under ncu is never a bench value; ncu stages run only after the plain command exited 0.
mkdir -p gpurun_out
PY=python
NCU="ncu --clock-control none"

summ() {  # one-line summary of a bench JSON line
  $PY - "$1" <<'EOF'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
except Exception as e:
    print("  (no JSON line:", e, ")"); sys.exit(0)
r = d.get("roofline") or {}
print("  value %.4g %s  %.3f ms/step  roofline %.1f %s frac %.3f  parity %s  clocks %s" % (
    d.get("value") or 0, d.get("unit"), d.get("ms_per_step") or 0, r.get("achieved") or 0, r.get("unit"), r.get("frac") or 0,
    (d.get("parity") or {}).get("ok"), (d.get("clocks") or {}).get("reasons")))
e = d.get("e2e") or {}
if e:
    print("  e2e %.4g %s" % (e.get("value") or 0, e.get("unit")), {k: (round(v.get("value", 0) / 1e9, 4) if "value" in v else v) for k, v in (e.get("variants") or {}).items()})
for k in ("collective", "strong"):
    if d.get(k): print(" ", k, json.dumps(d[k])[:400])
for s in d.get("secondary") or []:
    rr = s.get("roofline") or {}
    print("  secondary %-28s %.4g %s %.3f ms frac %.3f parity %s" % (s.get("workload"), s.get("value") or 0, s.get("unit"), s.get("ms_per_step") or 0, rr.get("frac") or 0, (s.get("parity") or {}).get("ok")))
EOF
}

stage_tests() {          # the whole -m gpu suite
  timeout 1500 $PY -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
}
stage_tests_new() {      # only the test files named in $TESTS
  timeout 900 $PY -m pytest $TESTS -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_new.log
}
stage_smoke() { timeout 300 $PY -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5; }
stage_bench() {          # default bench line (what the driver runs)
  timeout 1200 $PY bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err; summ gpurun_out/bench_n1.json
}
stage_bench_ref() {
  timeout 600 $PY bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_ref_n1.json
}
stage_workloads() {      # every other workload, kernel-only
  for wl in ${WORKLOADS:-cga_gp cga_gp_outer r3_gp sta_grade_inner_sum sta_fused_grade_inner_sum cl8_fixed_gp cl8_fixed_gp_f64 pga_points_compact}; do
    timeout 400 $PY bench.py --workload $wl --no-e2e --no-cpu --no-secondary --steps 10 > gpurun_out/wl_$wl.json 2> gpurun_out/wl_$wl.err
    echo "$wl rc=$?"; summ gpurun_out/wl_$wl.json
  done
}
stage_multi() {          # torchrun at $NGPU ranks: headline + collective + strong records
  N=${NGPU:-2}
  timeout 900 $PY -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 ${BENCH_ARGS} \
    > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n=$N rc=$?"; tail -3 gpurun_out/bench_n$N.err; summ gpurun_out/bench_n$N.json
}
stage_shard_check() {    # native single-process multi-GPU dispatcher against the oracle
  timeout 600 $PY tests/tools/multi_gpu_check.py 2>&1 | tail -20 | tee gpurun_out/shard_check.log
}
stage_shard_check_ranks() {  # the same check with one rank per GPU (ncclCommInitRank)
  N=${NGPU:-2}
  timeout 600 $PY -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/tools/multi_gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -12 | tee gpurun_out/shard_check_ranks.log
}
stage_peaks() {          # cuBLAS TF32 / FP64 / BF16 GEMM and copy micro-benchmarks -> gpurun_out/peaks_local.json
  timeout 300 $PY scripts/measure_peaks.py > gpurun_out/peaks_local.json 2> gpurun_out/peaks_local.err; echo "peaks rc=$?"; cat gpurun_out/peaks_local.json
}
stage_latency() {        # per-call latency of single Multivector operations
  timeout 300 $PY scripts/bench_single_latency.py 2>&1 | tee gpurun_out/single_latency.jsonl
}
stage_aux() { timeout 600 $PY scripts/bench_aux_ops.py > gpurun_out/aux_ops.jsonl 2> gpurun_out/aux_ops.err; echo "aux rc=$?"; cut -c1-200 gpurun_out/aux_ops.jsonl; }
stage_torch() { timeout 300 $PY scripts/bench_torch_interop.py > gpurun_out/torch_interop.jsonl 2> gpurun_out/torch_interop.err; echo "torch rc=$?"; cut -c1-300 gpurun_out/torch_interop.jsonl; }
stage_reftests() {       # the reference's own test files (vendored under tests/golden/reftests) against the drop-in module
  timeout 900 $PY -m pytest tests/test_reference_suites.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/reftests.log
}
stage_launches() {       # ncu launch list of the bench command (per-launch durations, cold cache, serialised)
  timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches.csv $PY bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu_launches.log 2>&1; echo "launches rc=$?"; grep -c sandwich gpurun_out/launches.csv
}
stage_ncu() {            # one --set full capture: NCU_KERNEL regex, NCU_CMD command, NCU_OUT name
  timeout 900 $NCU --set full --import-source on -k "regex:${NCU_KERNEL}" -c ${NCU_COUNT:-1} -s ${NCU_SKIP:-3} -o gpurun_out/${NCU_OUT} -f ${NCU_CMD} > gpurun_out/${NCU_OUT}.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/${NCU_OUT}.log
}
stage_fuzz() { for seed in ${SEEDS:-21 22}; do FUZZ_SEED=$seed timeout 120 $PY tests/tools/fuzz_parity.py 2000 > gpurun_out/fuzz_$seed.log 2>&1; echo "seed $seed rc=$?"; tail -2 gpurun_out/fuzz_$seed.log | cut -c1-300; done; }
stage_k8() {             # K8 (matrix-representation kernel): parity tests, then f32 / f64 at 2^24 rows per group count and debug mode
  timeout 600 $PY -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "block_diagonalised or tensor_cores or reference_layout" 2>&1 | tail -8 | tee gpurun_out/pytest_k8.log
  for g in ${K8_GROUPS:-3 2}; do for wl in cl8_fixed_gp_rep cl8_fixed_gp_f64; do
    LCC_K8_GROUPS=$g timeout 300 $PY bench.py --workload $wl --no-e2e --no-cpu --no-secondary --steps 10 > gpurun_out/k8_${wl}_g$g.json 2> gpurun_out/k8_${wl}_g$g.err
    echo "$wl groups=$g rc=$?"; summ gpurun_out/k8_${wl}_g$g.json
  done; done
  for dbg in ${K8_DEBUG:-1 2}; do for wl in cl8_fixed_gp_rep cl8_fixed_gp_f64; do
    LCC_K8_DEBUG=$dbg timeout 300 $PY bench.py --workload $wl --no-e2e --no-cpu --no-secondary --no-parity --steps 10 > gpurun_out/k8_${wl}_dbg$dbg.json 2> gpurun_out/k8_${wl}_dbg$dbg.err
    echo "$wl debug=$dbg rc=$?"; summ gpurun_out/k8_${wl}_dbg$dbg.json
  done; done
}
stage_k8_aos() {         # K8 on reference-layout units: parity, then f32 / f64 at 2^24 rows in both layouts (+ pass-through / transforms-only)
  timeout 900 $PY -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "reference_layout" 2>&1 | tail -8 | tee gpurun_out/pytest_k8_aos.log
  for lay in aos soa; do for wl in cl8_fixed_gp_rep cl8_fixed_gp_f64; do
    timeout 300 $PY bench.py --workload $wl --layout $lay --no-e2e --no-cpu --no-secondary --steps 10 > gpurun_out/k8_${wl}_$lay.json 2> gpurun_out/k8_${wl}_$lay.err
    echo "$wl $lay rc=$?"; summ gpurun_out/k8_${wl}_$lay.json
  done; done
  for dbg in ${K8_DEBUG-1 2}; do for wl in cl8_fixed_gp_rep cl8_fixed_gp_f64; do
    LCC_K8_DEBUG=$dbg timeout 300 $PY bench.py --workload $wl --layout aos --no-e2e --no-cpu --no-secondary --no-parity --steps 10 > gpurun_out/k8_${wl}_aos_dbg$dbg.json 2> gpurun_out/k8_${wl}_aos_dbg$dbg.err
    echo "$wl aos debug=$dbg rc=$?"; summ gpurun_out/k8_${wl}_aos_dbg$dbg.json
  done; done
  for lay in aos soa; do for wl in cl8_fixed_gp_rep cl8_fixed_gp_f64; do
    LCC_K8_GROUPS=2 timeout 300 $PY bench.py --workload $wl --layout $lay --no-e2e --no-cpu --no-secondary --steps 10 > gpurun_out/k8_${wl}_${lay}_g2.json 2> gpurun_out/k8_${wl}_${lay}_g2.err
    echo "$wl $lay groups=2 rc=$?"; summ gpurun_out/k8_${wl}_${lay}_g2.json
  done; done
}
stage_k8_mem() {         # K8 memory pipeline experiments: L2 promotion x unit order (pass-through and full), swizzle off (pass-through)
  run1() { # name, env..., workload, extra flags
    local name=$1; shift; local wl=$1; shift
    env "$@" timeout 300 $PY bench.py --workload $wl --no-e2e --no-cpu --no-secondary --steps 10 ${NOPAR} > gpurun_out/k8m_${wl}_$name.json 2> gpurun_out/k8m_${wl}_$name.err
    echo "$wl $name rc=$?"; summ gpurun_out/k8m_${wl}_$name.json
  }
  for wl in cl8_fixed_gp_rep cl8_fixed_gp_f64; do
    for promo in 128 256; do for order in 0 1; do
      NOPAR=--no-parity run1 pass_p${promo}_o${order} $wl LCC_K8_DEBUG=1 LCC_K8_L2PROMO=$promo LCC_K8_ORDER=$order
    done; done
    NOPAR=--no-parity run1 pass_noswz_p256_o1 $wl LCC_K8_DEBUG=1 LCC_K8_SWIZZLE=0
    NOPAR= run1 full_default $wl LCC_K8_GROUPS=3
    NOPAR= run1 full_default_g2 $wl LCC_K8_GROUPS=2
  done
}
stage_tma_probe() {      # which tile shape / unit scheduling lets a persistent TMA ring stream a 256-blade batch at the HBM rate
  mkdir -p build && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/tma_probe scripts/tma_passthrough_probe.cu \
    && timeout 300 build/tma_probe | tee gpurun_out/tma_probe.jsonl
}
stage_info() { nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit,memory.total --format=csv; nproc; free -g | head -2; nvidia-smi topo -m 2>/dev/null | head -12; }

for st in "$@"; do
  echo "=== stage $st"
  "stage_$st"
done
