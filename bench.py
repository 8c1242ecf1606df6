#!/usr/bin/env python3
"""bench.py -- headline benchmark of the batched Clifford-algebra hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

Default workload = BASELINE.json configs[1]: PGA Cl(3,0,1) motor sandwich applied to 2^28 points,
f32, dense 16-blade rows (128 algorithmic bytes per point: 64 read + 64 written).  A "step" is one
pass of the sandwich kernel over the whole device-resident batch.  For N > 1 (torchrun, one process
per GPU) every rank owns its own 2^28-point shard (weak scaling, no data-path collective); timing is
bracketed by a barrier + synchronize and the MAX over ranks is reported.  Kernels are launched on the
compute stream of the native shard group (csrc/shard_group.cu), CUDA events are recorded on that stream.

One JSON line on stdout (rank 0):
  value       points/s with inputs already resident in HBM (mean over K steps; ms_min / ms_median beside it)
  e2e         the same metric from HOST arrays through the drop-in Python API -- value = the reference's own call
              sequence MultivectorBatch.from_numpy(alg, x).sandwich(M).to_numpy() on a pinned NumPy array, H2D and D2H
              inside the timed region; `variants` adds the pageable-array run, the one-call MultivectorBatch.sandwich_numpy
              and the C-ABI entry point lcc_host_sandwich behind it, plus the same points as compact (N, 3) xyz rows
              (lcc_host_pga_points_sandwich: 24 bytes per point each way; a different data format, listed for comparison)
  roofline    algorithmic bytes / kernel time against the measured HBM peak (MEASURED_PEAKS.json); tensor-bound
              workloads against profiles/peaks_local.json (cuBLAS TF32 / FP64 measured with scripts/measure_peaks.py)
  cpu_baseline  the oracle port of the reference's Rust loop timed on this box's host cores (N = 1 only)
  collective  BASELINE config 4 fused (grade projection + inner product + batch sum in one kernel, 2^27 rows per GPU) with
              its NCCL all-reduce every step, through lcc_shard_product_sum_async; at N > 1 also the same kernel without the
              collective and on one GPU of the job alone
  strong      BASELINE config 2 with 2^28 points IN TOTAL (2^28 / N per GPU)
  secondary   the other BASELINE configs at full size on this rank's GPU (ms_per_step, roofline, parity, clocks each);
              config 5 (Cl(8,0,0) fixed-operand product) as K8 in both layouts and as the dense tensor-core GEMMs
`--impl reference` times only the CPU implementation (bounded sample per step) and prints the same
line shape with "impl": "reference" and the same `config` block.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (signature, dtype, log2 rows, algorithmic bytes per unit, unit)
    "pga_sandwich": ((3, 0, 1), "f32", 28, 128, "points/s"),
    "cga_gp": ((4, 1, 0), "f64", 26, 768, "products/s"),
    "cga_gp_outer": ((4, 1, 0), "f64", 26, 1024, "products/s"),
    "r3_gp": ((3, 0, 0), "f64", 20, 192, "products/s"),
    # grade(1) reads 4 blade rows and writes 16 (160 B), lc reads 2x16 and writes 16 (384 B), sum reads 16 (128 B)
    "sta_grade_inner_sum": ((1, 3, 0), "f64", 27, 160 + 384 + 128, "multivectors/s"),
}
# BASELINE config 4 fused: <x>_1 . w summed over the batch in ONE kernel: 4 + 16 blade rows read per multivector
WORKLOADS["sta_fused_grade_inner_sum"] = ((1, 3, 0), "f64", 27, (4 + 16) * 8, "multivectors/s")
# the same motor applied to COMPACT points, (N,3) xyz in and out, embed + sandwich + extract fused:
# 24 algorithmic bytes per point instead of 128.  Reported separately, never mixed into the headline.
WORKLOADS["pga_points_compact"] = ((3, 0, 1), "f32", 28, 24, "points/s")
# BASELINE config 5: fixed operand x batch in Cl(8,0,0) on the tensor cores (tf32 x3): bound = tensor
WORKLOADS["cl8_fixed_gp"] = ((8, 0, 0), "f32", 24, 2048, "products/s")
# the same product in the reference's own precision, on the FP64 tensor pipe (DMMA)
WORKLOADS["cl8_fixed_gp_f64"] = ((8, 0, 0), "f64", 24, 4096, "products/s")
# Library default for both precisions: K8, block-diagonalised in the M16(R) representation (cl8_fixed_gp_rep f32,
# cl8_fixed_gp_f64; HBM-side bound).  The dense nb x nb multiplication-matrix GEMMs are measured beside them (option
# matrix_rep = 0): cl8_fixed_gp = tcgen05 3xTF32 (tensor-bound), cl8_fixed_gp_f64_gemm = the round-1 DMMA GEMM
WORKLOADS["cl8_fixed_gp_rep"] = ((8, 0, 0), "f32", 24, 2048, "products/s")
WORKLOADS["cl8_fixed_gp_f64_gemm"] = ((8, 0, 0), "f64", 24, 4096, "products/s")
CL8_WORKLOADS = ("cl8_fixed_gp", "cl8_fixed_gp_f64", "cl8_fixed_gp_rep", "cl8_fixed_gp_f64_gemm")
# tensor flops of the dense multiplication-matrix GEMM (f32: three tf32 MMAs per term); the default K8 path needs 16x fewer
# flops and is HBM-bound, so only the *_gemm variants report against the tensor roofline
TENSOR_FLOPS_PER_UNIT = {"cl8_fixed_gp": 3 * 2 * 256 * 256, "cl8_fixed_gp_f64_gemm": 2 * 256 * 256}
# the headline input is a PGA point (12 of its 16 blade rows are zeros); this variant streams 16 dense rows
WORKLOADS["pga_sandwich_dense16"] = ((3, 0, 1), "f32", 28, 128, "points/s")
HEADLINE = "pga_sandwich"


TRAFFIC_SOURCE = ("static: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel at this size "
                  "(profiles/traffic.json names the capture); NOT measured in this run -- a bench process cannot read DRAM counters")


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# --------------------------------------------------------------------------------------------- motor (BASELINE cfg2)
def headline_motor():
    """M = T(4,5,6) * R(axis (1,2,3) through (0.5,-0.25,1), theta 0.7), built in f64 from the
    formulas of src/pga.rs:287-345,386-389 (the product T*R expanded by hand for a translator)."""
    axis = np.array([1.0, 2.0, 3.0])
    axis /= np.linalg.norm(axis)
    p = np.array([0.5, -0.25, 1.0])
    c, s = np.cos(0.35), np.sin(0.35)
    R = np.zeros(16)
    R[0], R[6], R[5], R[3] = c, -s * axis[0], s * axis[1], -s * axis[2]
    mom = np.cross(p, axis)
    R[9], R[10], R[12] = s * mom[0], s * mom[1], s * mom[2]
    T = np.zeros(16)
    T[0], T[9], T[10], T[12] = 1.0, 2.0, 2.5, 3.0
    import oracle  # used here only to compose two constants; not on any timed path

    return oracle.OracleAlgebra(3, 0, 1).geometric_product(T[None, :], R[None, :])[0]


# --------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self._stop = index, [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self.nv:
            self.t.join(timeout=1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------- CPU reference arm
def cpu_rate(workload, sample_rows, threads, repeats=1):
    """Oracle port of the reference loop on `sample_rows` rows of the same synthetic workload."""
    import oracle

    sig, dt, _, _, _ = WORKLOADS[workload]
    o = oracle.OracleAlgebra(*sig)
    ndt = np.float32 if dt == "f32" else np.float64
    rng = np.random.default_rng(42)
    nb = o.num_blades
    if workload == "pga_sandwich_dense16":
        x = rng.standard_normal((sample_rows, 16)).astype(ndt)
        motor = headline_motor()
        fn = lambda: o.sandwich(motor, x, threads=threads)  # noqa: E731
    elif workload == "pga_sandwich":
        pts = rng.standard_normal((sample_rows, 3)).astype(ndt)
        x = np.zeros((sample_rows, 16), ndt)
        x[:, 7], x[:, 14], x[:, 13], x[:, 11] = 1.0, -pts[:, 0], pts[:, 1], -pts[:, 2]
        motor = headline_motor()
        fn = lambda: o.sandwich(motor, x, threads=threads)  # noqa: E731
    else:
        a = rng.standard_normal((sample_rows, nb)).astype(ndt)
        b = rng.standard_normal((sample_rows, nb)).astype(ndt)
        if workload in CL8_WORKLOADS:
            fn = lambda: o.geometric_product(a[:1], b, threads=threads)  # noqa: E731
        elif workload == "cga_gp_outer":
            fn = lambda: (o.geometric_product(a, b, threads=threads), o.outer_product(a, b, threads=threads))  # noqa: E731
        elif workload in ("sta_grade_inner_sum", "sta_fused_grade_inner_sum"):
            fn = lambda: o.sum(o.left_contraction(o.grade(a, 1), b, threads=threads))  # noqa: E731
        else:
            fn = lambda: o.geometric_product(a, b, threads=threads)  # noqa: E731
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn()
        dtm = time.perf_counter() - t0
        best = dtm if best is None else min(best, dtm)
    return sample_rows / best, best


def cpu_sample_rows(workload):
    return {"pga_sandwich": 1 << 25, "cga_gp": 1 << 22, "cga_gp_outer": 1 << 22, "r3_gp": 1 << 20, "sta_grade_inner_sum": 1 << 23,
            "sta_fused_grade_inner_sum": 1 << 23,
            "cl8_fixed_gp": 1 << 14, "cl8_fixed_gp_f64": 1 << 14, "cl8_fixed_gp_rep": 1 << 14, "cl8_fixed_gp_f64_gemm": 1 << 14, "pga_points_compact": 1 << 25, "pga_sandwich_dense16": 1 << 25}[workload]


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port of the Rust
    loops -- the crate itself cannot be built here) on this box's host cores.  Prints the same `config`
    block as the GPU arm (same workload, rows per GPU); how the CPU run was bounded is under `cpu_baseline`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload
    sig, dt, log2n, bytes_per, unit = WORKLOADS[wl]
    threads = args.cpu_threads or 1
    rows = max(1 << 16, cpu_sample_rows(wl) // 4)  # per-step sample, bounded so K steps end within minutes
    for _ in range(min(args.warmup, 1)):
        cpu_rate(wl, 1 << 16, threads)
    times = []
    for _ in range(args.steps):
        _, t = cpu_rate(wl, rows, threads)
        times.append(t)
    ms = 1e3 * float(np.mean(times))
    value = rows / (ms / 1e3)
    cores = host_cores()  # torchrun exports OMP_NUM_THREADS=1: ask the scheduler, not OpenMP
    all_cores_value = cpu_rate(wl, rows, cores)[0] if cores > threads else value  # what an OpenMP-over-rows port would do
    n = 1 << (args.log2_rows or log2n)
    line = {
        "impl": "reference", "metric": f"{wl} throughput", "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dt, "data": "synthetic",
        "config": config_block(wl, n),
        "impl_config": {"layout": "AoS (N, num_blades) row-major in host memory (the reference's Array2)", "strict_fp": True,
                        "sample_rows_per_step": rows},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port",
                         "sample": f"{rows} rows per step of the same synthetic workload (the rate does not depend on the row count once out of "
                                   "cache); the reference is single-threaded (GIL held, no threads)",
                         "host_threads_available": cores, "all_cores_value": all_cores_value, "all_cores": cores},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)
    return 0


def workload_name(wl, n=None):
    sig, dt, log2n, _, _ = WORKLOADS[wl]
    desc = {
        "pga_sandwich": "PGA Cl(3,0,1) motor sandwich", "cga_gp": "CGA Cl(4,1,0) batched geometric product",
        "cga_gp_outer": "CGA Cl(4,1,0) fused geometric + outer product", "r3_gp": "Cl(3,0,0) batched geometric product",
        "sta_grade_inner_sum": "STA Cl(1,3,0) grade projection + inner product + batch sum",
        "cl8_fixed_gp": "Cl(8,0,0) fixed-operand left geometric product as the dense 256 x 256 GEMM (tcgen05, 3xTF32)",
        "cl8_fixed_gp_f64": "Cl(8,0,0) fixed-operand left geometric product, f64 (K8: block-diagonalised in the M16(R) representation)",
        "cl8_fixed_gp_rep": "Cl(8,0,0) fixed-operand left geometric product, f32 through K8 (block-diagonalised in the M16(R) representation)",
        "cl8_fixed_gp_f64_gemm": "Cl(8,0,0) fixed-operand left geometric product as the dense 256 x 256 GEMM, f64 (mma.sync m8n8k4 DMMA)",
        "sta_fused_grade_inner_sum": "STA Cl(1,3,0) fused grade projection + inner product + batch sum (one kernel)",
        "pga_points_compact": "PGA Cl(3,0,1) motor applied to compact (N,3) points, embed + sandwich + extract fused",
        "pga_sandwich_dense16": "PGA Cl(3,0,1) motor sandwich on fully dense 16-blade multivectors",
    }[wl]
    rows = f"2^{log2n}" if n is None or n == (1 << log2n) else (f"2^{n.bit_length() - 1}" if n & (n - 1) == 0 else str(n))
    return f"{desc}, {rows} rows per GPU, {dt}, {1 << sum(sig)} blades"


# --------------------------------------------------------------------------------------------- GPU arm
class Ctx:
    """Everything one rank needs to run workloads: library, device, the launching stream, the native shard group."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        from largecrimsoncanine_b200 import cabi, sharded

        self.torch, self.cabi, self.args = torch, cabi, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the batched path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            # the only collective of the path is a 16..256-element all-reduce: one channel (one CTA) is enough, and every
            # extra NCCL CTA spin-waiting for a slower rank takes SM slots away from the streaming kernel it overlaps with
            os.environ.setdefault("NCCL_MAX_NCHANNELS", "1")
            os.environ.setdefault("NCCL_MAX_CTAS", "1")
            dist.init_process_group("nccl", device_id=self.dev)  # plumbing only: barriers and the max-over-ranks of the timings
            self.dist = dist
        self.lib = cabi.lib()
        cabi.check(self.lib.lcc_set_device(self.local))
        # the native shard group (csrc/shard_group.cu): its compute stream is the launching stream of every workload, its
        # NCCL communicator carries the one collective of the path.  torch events are recorded on that same stream.
        self.group = sharded.ShardGroup.from_env(device=self.local)
        self.stream = torch.cuda.ExternalStream(self.group.stream(0), device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.sptr = C.c_void_p(self.group.stream(0))
        self.flags = cabi.LCC_FLAG_STRICT_FP if args.strict else 0
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(42 + self.rank)
        self._algs = {}

    def alg(self, sig):
        if sig not in self._algs:
            self._algs[sig] = self.cabi.Algebra(*sig)
        return self._algs[sig]

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if not self.dist:
            return v
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        self.group.close()
        if self.dist:
            self.dist.destroy_process_group()


def setup_workload(ctx, wl, n, layout="soa", collective=True, ld_pad=0):
    """Allocate the synthetic inputs of `wl` (n rows on this rank) and return (step, drain, parity, keepalive)."""
    torch, cabi, lib, dev, gen, sptr, flags = ctx.torch, ctx.cabi, ctx.lib, ctx.dev, ctx.gen, ctx.sptr, ctx.flags
    sig, dt, _, bytes_per, unit = WORKLOADS[wl]
    tdt = torch.float32 if dt == "f32" else torch.float64
    code = cabi.LCC_F32 if dt == "f32" else cabi.LCC_F64
    alg = ctx.alg(sig)
    nb = alg.num_blades
    aos_layout = layout == "aos"
    if aos_layout and wl not in ("pga_sandwich", "pga_sandwich_dense16", "cga_gp", "r3_gp", "cga_gp_outer", "cl8_fixed_gp_rep", "cl8_fixed_gp_f64"):
        raise SystemExit("--layout aos is available for pga_sandwich, cga_gp, cga_gp_outer, r3_gp and the K8 workloads cl8_fixed_gp_rep / cl8_fixed_gp_f64")

    def soa(t):  # torch (nb, n) tensor == engine-native SoA view with ld = n
        if aos_layout:  # torch (n, nb) tensor == the reference's row-major layout (src/batch.rs:41-47)
            return cabi.view(t.data_ptr(), t.shape[0], t.stride(0), code, cabi.LCC_AOS)
        return cabi.view(t.data_ptr(), t.shape[1], t.stride(0), code, cabi.LCC_SOA)

    def soa_alloc(fill_random):  # (nb, n) blade-major view whose blade rows are n + ld_pad elements apart
        t = torch.empty((nb, n + ld_pad), device=dev, dtype=tdt)
        if fill_random:
            t.normal_(generator=gen)
        else:
            t.zero_()
        return t[:, :n]

    def randn_batch():
        t = soa_alloc(True)
        return t.T.contiguous() if aos_layout else t

    def rows_of(t, idx):  # (len(idx), nb) host array of the selected rows
        return (t[idx] if aos_layout else t[:, idx].T).contiguous().cpu().numpy()

    import oracle

    o = oracle.OracleAlgebra(*sig)
    span = min(4096, max(1, n // 3))
    idx = torch.cat([torch.arange(0, span), torch.arange(n // 2, n // 2 + span), torch.arange(n - span, n)]).to(dev)
    tol = 1e-5 if dt == "f32" else 1e-12
    keep = {}

    def drain():
        pass

    def finish(got, want, rows):
        err = float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max() / np.abs(want).max())
        return {"rows_checked": int(rows), "max_rel_err": err, "tolerance": tol, "ok": bool(err <= tol)}

    if wl in ("pga_sandwich", "pga_sandwich_dense16"):
        if wl == "pga_sandwich":
            x = soa_alloc(False)
            pts = torch.randn((3, n), device=dev, dtype=tdt, generator=gen)
            x[7].fill_(1.0)
            x[14].copy_(-pts[0]); x[13].copy_(pts[1]); x[11].copy_(-pts[2])
            del pts
        else:  # every one of the 16 blade rows is dense random data (SURVEY 8d asks for both variants)
            x = soa_alloc(True)
        if aos_layout:
            x = x.T.contiguous()
        out = torch.empty_like(x) if aos_layout else soa_alloc(False)
        motor = headline_motor()
        mptr = motor.ctypes.data_as(C.POINTER(C.c_double))
        vx, vo = soa(x), soa(out)
        keep.update(x=x, out=out, motor=motor)

        def step():
            cabi.check(lib.lcc_batch_sandwich(alg.h, mptr, C.byref(vx), C.byref(vo), flags, sptr))

        def parity():
            return finish(rows_of(out, idx), o.sandwich(motor, rows_of(x, idx)), idx.numel())
    elif wl == "pga_points_compact":
        x = torch.randn((n, 3), device=dev, dtype=tdt, generator=gen)
        out = torch.empty_like(x)
        motor = headline_motor()
        mptr = motor.ctypes.data_as(C.POINTER(C.c_double))
        keep.update(x=x, out=out, motor=motor)

        def step():
            cabi.check(lib.lcc_pga_points_sandwich(alg.h, code, mptr, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), n, flags, sptr))

        def parity():
            return finish(out[idx].cpu().numpy(), oracle.pga_transform_points(motor, x[idx].cpu().numpy()), idx.numel())
    elif wl in ("cga_gp", "r3_gp", "cga_gp_outer"):
        a = randn_batch()
        b = randn_batch()
        out = torch.empty_like(a) if aos_layout else soa_alloc(False)
        va, vb, vo = soa(a), soa(b), soa(out)
        keep.update(a=a, b=b, out=out)
        if wl == "cga_gp_outer":
            out2 = torch.empty_like(a) if aos_layout else soa_alloc(False)
            vo2 = soa(out2)
            keep.update(out2=out2)

            def step():
                cabi.check(lib.lcc_batch_gp_outer(alg.h, C.byref(va), C.byref(vb), C.byref(vo), C.byref(vo2), flags, sptr))
        else:
            def step():
                cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(va), C.byref(vb), C.byref(vo), flags, sptr))

        def parity():
            xs, ys = rows_of(a, idx), rows_of(b, idx)
            res = finish(rows_of(out, idx), o.geometric_product(xs, ys), idx.numel())
            if wl == "cga_gp_outer":
                second = finish(rows_of(out2, idx), o.outer_product(xs, ys), idx.numel())
                res["max_rel_err"] = max(res["max_rel_err"], second["max_rel_err"])
                res["ok"] = res["ok"] and second["ok"]
            return res
    elif wl == "sta_fused_grade_inner_sum":
        a = torch.randn((nb, n), device=dev, dtype=tdt, generator=gen)
        b = torch.randn((nb, n), device=dev, dtype=tdt, generator=gen)
        keep.update(a=a, b=b)
        grp = ctx.group
        views_a, views_b = grp.views(), grp.views()
        views_a[0], views_b[0] = soa(a), soa(b)
        turn = [0]
        # BASELINE config 4, fused: per step ONE kernel reads <x>_1 and w and leaves 16 partial sums in a slot of the
        # group's ring; the NCCL all-reduce of that slot is enqueued behind it on the group's high-priority
        # communication stream (lcc_shard_product_sum_async), so the collective of step i overlaps the kernel of step
        # i+1.  Every all-reduce is fetched (drain) inside the timed region.  collective=False: the same kernel alone.
        if collective:
            def step():
                cabi.check(lib.lcc_shard_product_sum_async(grp.h, alg.h, cabi.LCC_OP_LC, views_a, 1, views_b, 0, turn[0] % 8, flags))
                turn[0] += 1

            res_host = np.empty(nb, np.float32 if dt == "f32" else np.float64)

            def drain():
                for slot in range(min(turn[0], 8)):
                    cabi.check(lib.lcc_shard_reduce_fetch(grp.h, alg.h, code, slot, res_host.ctypes.data))
        else:
            sums = torch.zeros(nb, device=dev, dtype=tdt)
            keep.update(sums=sums)

            def step():
                cabi.check(lib.lcc_batch_product_sum(alg.h, cabi.LCC_OP_LC, C.byref(views_a[0]), 1, C.byref(views_b[0]), C.c_void_p(sums.data_ptr()), flags, sptr))

        def parity():
            m = min(n, 1 << 20)  # the fused kernel only returns sums: re-run it on the first 2^20 rows and compare with the oracle
            sa, sb = cabi.view(a.data_ptr(), m, a.stride(0), code, cabi.LCC_SOA), cabi.view(b.data_ptr(), m, b.stride(0), code, cabi.LCC_SOA)
            chk = torch.zeros(nb, device=dev, dtype=tdt)
            cabi.check(lib.lcc_batch_product_sum(alg.h, cabi.LCC_OP_LC, C.byref(sa), 1, C.byref(sb), C.c_void_p(chk.data_ptr()), flags, sptr))
            torch.cuda.synchronize()
            xs, ys = a[:, :m].T.contiguous().cpu().numpy(), b[:, :m].T.contiguous().cpu().numpy()
            want, mag = o.sum(o.left_contraction(o.grade(xs, 1), ys))
            scale = np.abs(mag).max() / np.maximum(mag, 1e-300)  # compare on the scale of the summed magnitudes
            return finish(chk.cpu().numpy() * scale, want * scale, m)
    elif wl in CL8_WORKLOADS:
        # blade-major (nb, n) or, with --layout aos, the reference's (n, nb) rows: K8 then works on units of whole rows
        b = torch.randn((n, nb) if aos_layout else (nb, n), device=dev, dtype=tdt, generator=gen)
        a = torch.randn((nb, 4), device=dev, dtype=tdt, generator=gen)  # the fixed operand: one-row SoA view, ld = 4
        out = torch.empty_like(b)
        if aos_layout:
            vb, vo = (cabi.view(t.data_ptr(), n, nb, code, cabi.LCC_AOS) for t in (b, out))
        else:
            vb, vo = soa(b), soa(out)
        fixed = a[:, 0].to(torch.float64).cpu().numpy().copy()  # the fixed operand, known on the host (as in MultivectorBatch.from_numpy(1 x nb))
        fptr = fixed.ctypes.data_as(C.POINTER(C.c_double))
        keep.update(a=a, b=b, out=out, fixed=fixed)
        # lcc_batch_fixed_product; the option selects K8 (block-diagonalised in the 16 x 16 matrix representation) or the dense
        # multiplication-matrix GEMM on the tensor cores (K6 tcgen05 / K6d DMMA) -- see WORKLOADS for which is the default
        rep = {"cl8_fixed_gp": 0, "cl8_fixed_gp_f64_gemm": 0, "cl8_fixed_gp_rep": 2, "cl8_fixed_gp_f64": 2}[wl]

        def step():
            cabi.check(lib.lcc_set_option(b"matrix_rep", rep))
            cabi.check(lib.lcc_batch_fixed_product(alg.h, cabi.LCC_OP_GP, 0, fptr, C.byref(vb), C.byref(vo), flags, sptr))
            cabi.check(lib.lcc_set_option(b"matrix_rep", 2))  # back to the library default

        def parity():
            sub = idx[::64]
            ys = (b[sub] if aos_layout else b[:, sub].T).contiguous().cpu().numpy().astype(np.float64)
            want = o.geometric_product(a[:, 0].cpu().numpy().astype(np.float64)[None, :], ys)
            return finish((out[sub] if aos_layout else out[:, sub].T).contiguous().cpu().numpy(), want, sub.numel())
    elif wl == "sta_grade_inner_sum":
        a = torch.randn((nb, n), device=dev, dtype=tdt, generator=gen)
        b = torch.randn((nb, n), device=dev, dtype=tdt, generator=gen)
        g1 = torch.empty_like(a)
        out = torch.empty_like(a)
        sums = torch.zeros(nb, device=dev, dtype=tdt)
        va, vb, vg, vo = soa(a), soa(b), soa(g1), soa(out)
        keep.update(a=a, b=b, g1=g1, out=out, sums=sums)

        def step():
            cabi.check(lib.lcc_batch_unary(alg.h, cabi.LCC_UN_GRADE, 1, C.byref(va), C.byref(vg), sptr))
            cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_LC, C.byref(vg), C.byref(vb), C.byref(vo), flags, sptr))
            cabi.check(lib.lcc_batch_sum(alg.h, C.byref(vo), C.c_void_p(sums.data_ptr()), sptr))

        def parity():
            xs, ys = a[:, idx].T.contiguous().cpu().numpy(), b[:, idx].T.contiguous().cpu().numpy()
            return finish(out[:, idx].T.contiguous().cpu().numpy(), o.left_contraction(o.grade(xs, 1), ys), idx.numel())
    else:
        raise SystemExit(f"unknown workload {wl}")
    return step, drain, parity, keep


def measure(ctx, wl, n, steps, warmup, layout="soa", collective=True, parity=True, solo=False, ld_pad=0, graph=False):
    """W warm-up steps, then exactly K timed steps bracketed by barrier + synchronize; CUDA events on the launching
    stream around every step (mean = the contract's ms_per_step, min / median beside it); max over ranks.
    solo=True: only rank 0 runs (the other ranks idle at the barriers) -- the N = 1 figure measured inside an N > 1 job."""
    torch = ctx.torch
    active = (not solo) or ctx.rank == 0
    step, drain, parity_fn, keep = setup_workload(ctx, wl, n, layout, collective, ld_pad) if active else (None, None, None, {})
    if active:
        for _ in range(max(warmup, 3)):
            step()
        drain()
    ctx.barrier()
    ctx.lib.lcc_kernel_launch_count_reset()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    with ClockSampler(ctx.local) as clocks:
        ctx.barrier()
        if active:
            ev[0].record(ctx.stream)
            for i in range(steps):
                step()
                if i + 1 < steps:
                    ev[i + 1].record(ctx.stream)
            drain()  # every outstanding all-reduce is inside the timed region
            ev[steps].record(ctx.stream)
        ctx.barrier()
    launches = int(ctx.lib.lcc_kernel_launch_count())
    if active:
        total_ms = ev[0].elapsed_time(ev[steps])
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    else:
        total_ms, per = 0.0, [0.0]
    if not solo:
        total_ms = ctx.max_over_ranks(total_ms)
    rec = {"ms_per_step": total_ms / steps, "ms_min": float(np.min(per)), "ms_median": float(np.median(per)), "steps": steps,
           "gpu_launches": launches, "clocks": clocks.summary()}
    if graph and active:
        # launch-bound configurations: the same K steps captured once into a CUDA graph and replayed (no per-launch host
        # work, back-to-back kernel starts) -- reported beside the stream-launch figure, never instead of it
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=ctx.stream):
                for _ in range(steps):
                    step()
            g.replay()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(ctx.stream)
            for _ in range(5):
                g.replay()
            g1.record(ctx.stream)
            torch.cuda.synchronize()
            rec["ms_per_step_cuda_graph"] = g0.elapsed_time(g1) / (5 * steps)
        except Exception as exc:
            rec["ms_per_step_cuda_graph"] = None
            rec["cuda_graph_error"] = f"{type(exc).__name__}: {exc}"
    rec["parity"] = parity_fn() if (parity and active and ctx.rank == 0) else None
    return rec, keep


def release(ctx, keep):
    keep.clear()
    ctx.torch.cuda.synchronize()
    ctx.torch.cuda.empty_cache()


def traffic_of(wl):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(wl)
    except Exception:
        return None


def config_block(wl, n):
    """The workload definition -- identical for the GPU arm and the --impl reference arm (same_config)."""
    _, _, log2n, bytes_per, _ = WORKLOADS[wl]
    return {"workload": workload_name(wl, n), "rows_per_gpu": n, "l2_policy": l2_policy(bytes_per * n)}


def secondary_record(ctx, wl, steps, peaks, peak_kind, n=None, layout="soa"):
    sig, dt, log2n, bytes_per, unit = WORKLOADS[wl]
    n = n or (1 << log2n)
    rec, keep = measure(ctx, wl, n, steps, 3, layout, graph=(wl == "r3_gp"))
    release(ctx, keep)
    if ctx.rank != 0:
        return None
    out = {"workload": wl, "config": config_block(wl, n), "layout": layout, "dtype": dt, "unit": unit, "n_gpus": ctx.world,
           "value": ctx.world * n / (rec["ms_per_step"] / 1e3),
           "roofline": roofline_block(wl, n, rec["ms_per_step"], bytes_per, peaks, peak_kind, traffic_of(wl))}
    out.update(rec)
    if rec.get("ms_per_step_cuda_graph"):
        out["roofline_cuda_graph"] = roofline_block(wl, n, rec["ms_per_step_cuda_graph"], bytes_per, peaks, peak_kind, None)
    return out


def run_gpu_arm(args):
    ctx = Ctx(args)
    torch, cabi, lib, rank, world = ctx.torch, ctx.cabi, ctx.lib, ctx.rank, ctx.world
    wl = args.workload
    sig, dt, log2n, bytes_per, unit = WORKLOADS[wl]
    if args.log2_rows:
        log2n = args.log2_rows
    n = 1 << log2n
    code = cabi.LCC_F32 if dt == "f32" else cabi.LCC_F64
    alg = ctx.alg(sig)
    nb = alg.num_blades
    peaks, peak_kind = load_peaks()

    rec, keep = measure(ctx, wl, n, args.steps, args.warmup, args.layout, parity=not args.no_parity, ld_pad=args.ld_pad)
    ms_per_step = rec["ms_per_step"]
    value = world * n / (ms_per_step / 1e3)
    motor = keep.get("motor")
    release(ctx, keep)

    # ---- end to end through the drop-in Python API / the C-ABI host-buffer entry points (headline workload only)
    e2e = None
    if wl == "pga_sandwich" and not args.no_e2e:
        e2e = measure_e2e(args, lib, alg, code, motor, n, nb, world, rank, ctx.dev, ctx.dist, ctx.flags)
    elif wl == "pga_points_compact" and not args.no_e2e:
        e2e = measure_e2e(args, lib, alg, code, motor, n, nb, world, rank, ctx.dev, ctx.dist, ctx.flags, compact=True)
    if e2e is not None:
        import largecrimsoncanine as L

        L.host_cache_trim()

    # ---- the default line also carries: the collective-bearing config (4, fused) with its all-reduce, config 2 sharded
    # strongly (2^28 points in total), and the other BASELINE configs at full size on this rank's GPU
    collective = strong = None
    secondary = []
    if wl == HEADLINE and not args.no_secondary and not args.log2_rows:
        k = max(3, min(args.steps, 10))
        cw = "sta_fused_grade_inner_sum"
        cn = 1 << WORKLOADS[cw][2]
        crec, ckeep = measure(ctx, cw, cn, k, 3, collective=True)
        release(ctx, ckeep)
        collective = {"workload": workload_name(cw, cn), "rows_per_gpu": cn, "n_gpus": world, "scaling": "weak",
                      "api": "lcc_shard_product_sum_async + lcc_shard_reduce_fetch (native group: fused partial -> ncclAllReduce(16 x f64) "
                             "on a high-priority stream, every all-reduce inside the timed region)",
                      "value": world * cn / (crec["ms_per_step"] / 1e3), "unit": WORKLOADS[cw][4], "nccl_version": int(lib.lcc_shard_nccl_version())}
        collective.update(crec)
        if world > 1:
            nrec, nkeep = measure(ctx, cw, cn, k, 3, collective=False, parity=False)
            release(ctx, nkeep)
            srec, skeep = measure(ctx, cw, cn, k, 3, collective=False, parity=False, solo=True)
            release(ctx, skeep)
            collective.update(ms_per_step_without_collective=nrec["ms_per_step"], ms_per_step_one_gpu_alone=srec["ms_per_step"],
                              efficiency_vs_one_gpu_alone=srec["ms_per_step"] / crec["ms_per_step"] if rank == 0 else None,
                              note="one_gpu_alone: rank 0 runs the same kernel while the other GPUs of the box idle, measured in this same job")
        sn = (1 << WORKLOADS[HEADLINE][2]) // world
        srec2, skeep2 = measure(ctx, HEADLINE, sn, k, 3, parity=True)
        release(ctx, skeep2)
        strong = {"workload": workload_name(HEADLINE, sn), "rows_total": sn * world, "rows_per_gpu": sn, "n_gpus": world, "scaling": "strong",
                  "value": world * sn / (srec2["ms_per_step"] / 1e3), "unit": unit,
                  "roofline_per_gpu": roofline_block(HEADLINE, sn, srec2["ms_per_step"], bytes_per, peaks, peak_kind, None)}
        strong.update(srec2)
        for swl, slayout in (("r3_gp", "soa"), ("cga_gp_outer", "soa"), ("cga_gp", "aos"), ("sta_grade_inner_sum", "soa"), ("cl8_fixed_gp", "soa"),
                             ("cl8_fixed_gp_f64", "soa"), ("cl8_fixed_gp_f64", "aos"), ("cl8_fixed_gp_rep", "soa"), ("cl8_fixed_gp_rep", "aos"), ("cl8_fixed_gp_f64_gemm", "soa"), ("pga_sandwich_dense16", "soa"), ("pga_sandwich", "aos")):
            try:
                r = secondary_record(ctx, swl, k, peaks, peak_kind, layout=slayout)
            except Exception as exc:  # one workload failing must not take the headline line down
                r = {"workload": swl, "error": f"{type(exc).__name__}: {exc}"} if rank == 0 else None
            if r:
                secondary.append(r)

    if rank != 0:
        ctx.close()
        return 0

    cpu = None
    if not args.no_cpu and world == 1:  # CPU baseline on rank 0 at N = 1 only
        rows = cpu_sample_rows(wl)
        rate1, t1 = cpu_rate(wl, rows, 1)
        cores = host_cores()
        rate_all, _ = cpu_rate(wl, rows, cores) if cores > 1 else (rate1, t1)
        cpu = {"value": rate1, "unit": unit, "cores": 1, "kind": "port",
               "sample": f"{rows} rows of the same synthetic workload ({t1:.1f} s); the reference is single-threaded",
               "all_cores_value": rate_all, "all_cores": cores}
    aos_layout = args.layout == "aos"
    line = {
        "metric": f"{wl} throughput", "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "ms_min": rec["ms_min"], "ms_median": rec["ms_median"],
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dt, "data": "synthetic",
        "config": config_block(wl, n),
        "impl_config": {"layout": "dense (N,3) xyz rows" if wl == "pga_points_compact" else ("AoS (N, num_blades) row-major in HBM, TMA-staged tiles" if aos_layout else "SoA (blade-major) in HBM"),
                        "strict_fp": bool(args.strict), "launch_stream": "compute stream of the native shard group (lcc_shard_group_stream)"},
        "roofline": roofline_block(wl, n, ms_per_step, bytes_per, peaks, peak_kind, traffic_of(wl)),
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": rec["gpu_launches"], "clocks": rec["clocks"], "parity": rec["parity"],
        "collective": collective, "strong": strong, "secondary": secondary,
    }
    emit_line(line)
    ctx.close()
    return 0


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def l2_policy(bytes_per_pass):
    gb = bytes_per_pass / 1e9
    if bytes_per_pass >= 8 * 126e6:
        return f"no flush needed: every step streams {gb:.1f} GB of distinct inputs and outputs through a 126 MB L2"
    return (f"NOT L2-cold: a step touches only {gb:.3f} GB, comparable to the 126 MB L2 -- this configuration is the reference's "
            "CPU-runnable case and is reported for completeness; use --log2-rows 26 for the HBM-bound figure")


def local_peaks():
    """profiles/peaks_local.json: cuBLAS TF32 / FP64 GEMM peaks micro-benchmarked on a B200 of this pool with
    scripts/measure_peaks.py (MEASURED_PEAKS.json carries only the bf16 figure and the copy bandwidth)."""
    try:
        with open(os.path.join(ROOT, "profiles", "peaks_local.json")) as f:
            return json.load(f)
    except Exception:
        return None


def roofline_block(wl, n, kernel_ms, bytes_per, peaks, peak_kind, traffic):
    hbm = bytes_per * n / (kernel_ms / 1e3) / 1e9
    hbm_peak = float(peaks["hbm_gbs"])
    if wl in TENSOR_FLOPS_PER_UNIT:
        # dense contraction.  `achieved` counts the flops of the multiplication-matrix GEMM the north_star names
        # (2 * nb^2 per row; x3 for the 3xTF32 split), whatever algorithm the kernel actually runs; the HBM figures beside
        # it say how far the kernel is from the streaming floor of the same rows.
        tf = TENSOR_FLOPS_PER_UNIT[wl] * n / (kernel_ms / 1e3) / 1e12
        lp = local_peaks()
        key = "fp64" if wl == "cl8_fixed_gp_f64_gemm" else "tf32"
        if lp and key in lp:
            peak, sustained = float(lp[key]["burst_tflops"]), float(lp[key]["sustained_tflops"])
            src = f"profiles/peaks_local.json: cuBLAS {key} GEMM 8192^3 measured with scripts/measure_peaks.py (burst; sustained beside it)"
        elif key == "fp64":
            peak, sustained, src = 40.0, None, "nominal B200 FP64 tensor rate (40 TFLOP/s); no measured figure available"
        else:
            peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
            sustained = float(peaks.get("bf16_tflops_sustained", 0.0)) / 2.0 or None
            src = f"{peak_kind} bf16_tflops / 2 (tf32 runs at half the bf16 rate; no direct tf32 measurement available)"
        return {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": traffic,
                "traffic_source": TRAFFIC_SOURCE if traffic else None, "peak_source": src,
                "peak_sustained": sustained, "frac_of_sustained": (tf / sustained) if sustained else None,
                "tensor_flops_per_unit": TENSOR_FLOPS_PER_UNIT[wl], "useful_tflops": tf / 3.0 if key == "tf32" else tf,
                "hbm_gbs": hbm, "hbm_frac": hbm / hbm_peak, "algorithmic_bytes_per_unit": bytes_per}
    return {"bound": "hbm", "achieved": hbm, "peak": hbm_peak, "unit": "GB/s", "frac": hbm / hbm_peak, "traffic": traffic,
            "traffic_source": TRAFFIC_SOURCE if traffic else None,
            "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)", "algorithmic_bytes_per_unit": bytes_per,
            "frac_of_nominal_8TBs": hbm / 8000.0}


def bind_to_gpu_numa_node(local):
    """Pin this process to the CPUs NVML reports as local to its GPU, so the pinned host buffers are
    first-touched on the NUMA node the GPU's PCIe root hangs off (matters when 8 ranks share one host)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        return True
    except Exception:
        return False


def measure_e2e(args, lib, alg, code, motor, n, nb, world, rank, dev, dist, flags, compact=False):
    """End to end from HOST arrays in the reference's (N, 16) row-major layout, copies inside the timed region.

    value = the reference's own call sequence through the drop-in Python API,
            MultivectorBatch.from_numpy(alg, x).sandwich(M).to_numpy()            (src/batch.rs:79-94,434-503,233-235)
            on a pinned NumPy array.  Its upload and download cannot overlap (each call must finish before the next
            starts), so it is bounded by one PCIe direction at a time.
    variants: the same sequence on a PAGEABLE array; the fused one-call form MultivectorBatch.sandwich_numpy (both PCIe
            directions busy at once); and the C-ABI entry point lcc_host_sandwich it maps to.
    With `compact`, lcc_host_pga_points_sandwich on pinned (N, 3) xyz buffers (the pga_points_compact workload)."""
    import psutil
    import torch

    numa_bound = bind_to_gpu_numa_node(dev.index or 0) if world > 1 else False

    from largecrimsoncanine_b200 import cabi

    es = 4 if code == cabi.LCC_F32 else 8
    if compact:
        return measure_e2e_compact(args, lib, alg, code, motor, n, es, world, rank, dev, dist, flags, numa_bound)
    import largecrimsoncanine as L

    n_e2e = n
    copies = 3 if world == 1 else 2  # pinned in, pinned out (+ one pageable input at N = 1)
    need = copies * n_e2e * nb * es * max(world, 1)
    avail = psutil.virtual_memory().available
    while need > 0.5 * avail and n_e2e > (1 << 20):
        n_e2e //= 2
        need //= 2
    ndt = np.float32 if es == 4 else np.float64
    L.set_device(dev.index or 0)
    palg = L.Algebra(3, 0, 1)
    pmotor = L.Multivector.from_list_in([float(v) for v in motor], palg)
    xh = L.pinned_empty((n_e2e, nb), ndt)
    oh = L.pinned_empty((n_e2e, nb), ndt)
    g = np.random.default_rng(7 + rank)
    blk = 1 << 22
    for s in range(0, n_e2e, blk):  # fill in blocks: (N,3) normals embedded as PGA points
        e = min(n_e2e, s + blk)
        p = g.standard_normal((e - s, 3)).astype(ndt)
        xh[s:e] = 0
        xh[s:e, 7] = 1.0
        xh[s:e, 14], xh[s:e, 13], xh[s:e, 11] = -p[:, 0], p[:, 1], -p[:, 2]
    mptr = motor.ctypes.data_as(C.POINTER(C.c_double))
    if flags:
        L.set_strict_fp(True)

    def three_call(x):
        return L.MultivectorBatch.from_numpy(palg, x).sandwich(pmotor).to_numpy()

    def fused(x):
        return L.MultivectorBatch.sandwich_numpy(palg, pmotor, x, out=oh)

    def c_abi(x):
        cabi.check(lib.lcc_host_sandwich(alg.h, code, mptr, C.c_void_p(x.ctypes.data), C.c_void_p(oh.ctypes.data), n_e2e, flags))
        return oh

    import oracle

    o = oracle.OracleAlgebra(3, 0, 1)
    want_head = o.sandwich(motor, xh[:2048])
    want_tail = o.sandwich(motor, xh[n_e2e - 2048:])

    def timed(fn, x, steps):
        out = fn(x)  # warm-up: stream-ordered pools, pinned cache, staging buffers
        ok = bool(np.abs(out[:2048] - want_head).max() <= 1e-4 and np.abs(out[n_e2e - 2048:] - want_tail).max() <= 1e-4)
        del out
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = fn(x)
            del out
        torch.cuda.synchronize()
        dt_s = time.perf_counter() - t0
        if dist:
            t = torch.tensor([dt_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_s = float(t.item())
        return {"value": world * n_e2e * steps / dt_s, "s_per_step": dt_s / steps, "steps": steps, "checked_against_oracle": ok}

    steps = max(1, min(args.steps, args.e2e_steps))
    variants = {}
    plan = [("three_call_pinned", three_call, "pinned", steps), ("fused_pinned", fused, "pinned", steps), ("c_abi_pinned", c_abi, "pinned", steps)]
    if world == 1:
        plan += [("three_call_pageable", three_call, "pageable", max(1, steps // 2)), ("fused_pageable", fused, "pageable", max(1, steps // 2))]
    xp = None
    for name, fn, kind, k in plan:
        try:
            if kind == "pageable" and xp is None:
                xp = np.empty((n_e2e, nb), ndt)
                for s in range(0, n_e2e, blk):
                    xp[s:min(n_e2e, s + blk)] = xh[s:min(n_e2e, s + blk)]
            variants[name] = timed(fn, xh if kind == "pinned" else xp, k)
        except Exception as exc:  # a variant that cannot run (host memory) must not take the bench line down
            variants[name] = {"error": f"{type(exc).__name__}: {exc}"}
    try:
        # the same points and motor in compact form: (N, 3) xyz in and out, embed + sandwich + extract fused on the device
        # (24 bytes per point over PCIe instead of 256) -- a different data format from the metric's 16-blade rows, listed
        # for what the 75 % structural zeros of a PGA point cost end to end
        ch, co = L.pinned_empty((n_e2e, 3), ndt), L.pinned_empty((n_e2e, 3), ndt)
        for s in range(0, n_e2e, blk):
            e = min(n_e2e, s + blk)
            ch[s:e, 0], ch[s:e, 1], ch[s:e, 2] = -xh[s:e, 14], xh[s:e, 13], -xh[s:e, 11]

        def compact_call():
            cabi.check(lib.lcc_host_pga_points_sandwich(alg.h, code, mptr, C.c_void_p(ch.ctypes.data), C.c_void_p(co.ctypes.data), n_e2e, flags))

        compact_call()
        sample = np.r_[0:2048, n_e2e - 2048:n_e2e]
        ok = bool(np.abs(co[sample] - oracle.pga_transform_points(motor, ch[sample])).max() <= 1e-4)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            compact_call()
        torch.cuda.synchronize()
        dt_s = time.perf_counter() - t0
        if dist:
            t = torch.tensor([dt_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_s = float(t.item())
        variants["compact_xyz_c_abi_pinned"] = {"value": world * n_e2e * steps / dt_s, "s_per_step": dt_s / steps, "steps": steps, "checked_against_oracle": ok,
                                                "h2d_bytes_per_step": n_e2e * 3 * es, "d2h_bytes_per_step": n_e2e * 3 * es}
        del ch, co
    except Exception as exc:
        variants["compact_xyz_c_abi_pinned"] = {"error": f"{type(exc).__name__}: {exc}"}
    if flags:
        L.set_strict_fp(False)
    head = variants.get("three_call_pinned", {})
    bytes_step = n_e2e * nb * es
    return {"value": head.get("value"), "unit": "points/s", "h2d_bytes_per_step": bytes_step, "d2h_bytes_per_step": bytes_step,
            "points_per_gpu": n_e2e, "steps": head.get("steps"),
            "api": "largecrimsoncanine.MultivectorBatch.from_numpy(alg, x).sandwich(M).to_numpy() on a pinned NumPy array "
                   "(the reference's own call sequence; chunked lcc_batch_upload / kernel / lcc_batch_download)",
            "timer": "host wall clock around the blocking calls, max over ranks",
            "checked_against_oracle": head.get("checked_against_oracle"), "numa_bound": numa_bound,
            "host_copy_threads": int(os.environ.get("LCC_HOST_COPY_THREADS", "0")) or None,
            "variants": variants,
            "variants_note": "three_call_*: from_numpy -> sandwich -> to_numpy (upload and download are serial by the API's "
                             "semantics); fused_*: MultivectorBatch.sandwich_numpy (one call, both PCIe directions at once); "
                             "c_abi_pinned: lcc_host_sandwich, the C-ABI entry point behind sandwich_numpy; compact_xyz_c_abi_pinned: "
                             "lcc_host_pga_points_sandwich on the same points as (N, 3) xyz rows (24 B per point each way "
                             "instead of 128: not the metric's data format, listed for comparison)"}


def measure_e2e_compact(args, lib, alg, code, motor, n, es, world, rank, dev, dist, flags, numa_bound):
    import torch

    from largecrimsoncanine_b200 import cabi

    tdt = torch.float32 if es == 4 else torch.float64
    xh = torch.empty((n, 3), dtype=tdt, pin_memory=True)
    oh = torch.empty((n, 3), dtype=tdt, pin_memory=True)
    g = torch.Generator().manual_seed(7 + rank)
    blk = 1 << 24
    for s in range(0, n, blk):
        e = min(n, s + blk)
        xh[s:e].normal_(generator=g)
    mptr = motor.ctypes.data_as(C.POINTER(C.c_double))

    def call():
        cabi.check(lib.lcc_host_pga_points_sandwich(alg.h, code, mptr, C.c_void_p(xh.data_ptr()), C.c_void_p(oh.data_ptr()), n, flags))

    steps = max(1, min(args.steps, args.e2e_steps))
    call()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    torch.cuda.synchronize()
    dt_s = time.perf_counter() - t0
    if dist:
        t = torch.tensor([dt_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_s = float(t.item())
    import oracle

    sample = np.r_[0:2048, n - 2048:n]
    want = oracle.pga_transform_points(motor, xh[sample].numpy())
    ok = bool(np.abs(oh[sample].numpy() - want).max() <= 1e-4)
    return {"value": world * n * steps / dt_s, "unit": "points/s", "h2d_bytes_per_step": n * 3 * es, "d2h_bytes_per_step": n * 3 * es,
            "points_per_gpu": n, "steps": steps,
            "api": "lcc_host_pga_points_sandwich (C ABI, pinned host (N,3) buffers, chunked H2D/kernel/D2H pipeline)",
            "timer": "host wall clock around the blocking call, max over ranks", "checked_against_oracle": ok, "numa_bound": numa_bound}


_JSON_FD = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that writes to file descriptor 1 during the run (NCCL prints
    its version banner there when NCCL_DEBUG is set on the box) is sent to stderr instead."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_line(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--log2-rows", type=int, default=0, help="override rows per GPU (power of two exponent)")
    ap.add_argument("--ld-pad", type=int, default=0, help="experiment: extra elements between consecutive blade rows of the SoA operands (a multiple of 64)")
    ap.add_argument("--layout", default="soa", choices=["soa", "aos"], help="HBM layout of the batches (aos = the reference's (N, num_blades) rows)")
    ap.add_argument("--strict", action="store_true", help="LCC_FLAG_STRICT_FP (bit-exact, no FMA)")
    ap.add_argument("--cpu-threads", type=int, default=0, help="reference arm: OpenMP threads (default 1 = the reference's behaviour)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads / collective / strong records of the default line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
