#!/usr/bin/env python3
"""bench.py -- headline benchmark of the batched Clifford-algebra hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

Default workload = BASELINE.json configs[1]: PGA Cl(3,0,1) motor sandwich applied to 2^28 points,
f32, dense 16-blade rows (128 algorithmic bytes per point: 64 read + 64 written).  A "step" is one
pass of the sandwich kernel over the whole device-resident batch.  For N > 1 (torchrun, one process
per GPU) every rank owns its own 2^28-point shard (weak scaling, no data-path collective); timing is
bracketed by a barrier + synchronize and the MAX over ranks is reported.

One JSON line on stdout (rank 0):
  value       points/s with inputs already resident in HBM (CUDA events on the launching stream)
  e2e         the same metric through the C-ABI host-buffer entry point lcc_host_sandwich: pinned HOST
              buffers in the reference's (N, 16) layout, H2D and D2H copies inside the timed region
  roofline    algorithmic bytes / kernel time against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle port of the reference's Rust loop timed on this box's host cores
`--impl reference` times only that CPU implementation (bounded sample per step) and prints the same
line shape with "impl": "reference".
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
TENSOR_FLOPS_PER_UNIT = {"cl8_fixed_gp": 3 * 2 * 256 * 256, "cl8_fixed_gp_f64": 2 * 256 * 256}  # f32: three tf32 MMAs per term
HEADLINE = "pga_sandwich"


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
    if workload == "pga_sandwich":
        pts = rng.standard_normal((sample_rows, 3)).astype(ndt)
        x = np.zeros((sample_rows, 16), ndt)
        x[:, 7], x[:, 14], x[:, 13], x[:, 11] = 1.0, -pts[:, 0], pts[:, 1], -pts[:, 2]
        motor = headline_motor()
        fn = lambda: o.sandwich(motor, x, threads=threads)  # noqa: E731
    else:
        a = rng.standard_normal((sample_rows, nb)).astype(ndt)
        b = rng.standard_normal((sample_rows, nb)).astype(ndt)
        if workload in ("cl8_fixed_gp", "cl8_fixed_gp_f64"):
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
            "cl8_fixed_gp": 1 << 14, "cl8_fixed_gp_f64": 1 << 14, "pga_points_compact": 1 << 25}[workload]


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port of the Rust
    loops -- the crate itself cannot be built here) on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle

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
    cores = oracle.max_threads()
    all_cores_value = cpu_rate(wl, rows, cores)[0] if cores > threads else value  # what an OpenMP-over-rows port would do
    n = 1 << (args.log2_rows or log2n)
    line = {
        "impl": "reference", "metric": f"{wl} throughput", "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dt, "data": "synthetic",
        "config": {"workload": workload_name(wl), "rows_per_gpu": n, "layout": "AoS (N, num_blades) row-major in host memory (the reference's Array2)",
                   "l2_policy": "n/a (CPU arm)", "strict_fp": True, "sample_rows_per_step": rows},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port",
                         "sample": f"{rows} rows per step of the same synthetic workload; the reference is single-threaded (GIL held, no threads)",
                         "host_threads_available": cores, "all_cores_value": all_cores_value, "all_cores": cores},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)
    return 0


def workload_name(wl):
    sig, dt, log2n, _, _ = WORKLOADS[wl]
    desc = {
        "pga_sandwich": "PGA Cl(3,0,1) motor sandwich", "cga_gp": "CGA Cl(4,1,0) batched geometric product",
        "cga_gp_outer": "CGA Cl(4,1,0) fused geometric + outer product", "r3_gp": "Cl(3,0,0) batched geometric product",
        "sta_grade_inner_sum": "STA Cl(1,3,0) grade projection + inner product + batch sum",
        "cl8_fixed_gp": "Cl(8,0,0) fixed-operand left geometric product (tcgen05, 3xTF32)",
        "cl8_fixed_gp_f64": "Cl(8,0,0) fixed-operand left geometric product, f64 (mma.sync m8n8k4 DMMA)",
        "sta_fused_grade_inner_sum": "STA Cl(1,3,0) fused grade projection + inner product + batch sum (one kernel)",
        "pga_points_compact": "PGA Cl(3,0,1) motor applied to compact (N,3) points, embed + sandwich + extract fused",
    }[wl]
    return f"{desc}, 2^{log2n} rows per GPU, {dt}, {1 << sum(sig)} blades"


# --------------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from largecrimsoncanine_b200 import cabi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the batched path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the only collective of the path is a 16..256-element all-reduce: one channel (one CTA) is enough, and every
        # extra NCCL CTA spin-waiting for a slower rank takes SM slots away from the streaming kernel it overlaps with
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "1")
        os.environ.setdefault("NCCL_MAX_CTAS", "1")
        dist.init_process_group("nccl", device_id=dev)
    lib = cabi.lib()
    cabi.check(lib.lcc_set_device(local))

    wl = args.workload
    sig, dt, log2n, bytes_per, unit = WORKLOADS[wl]
    if args.log2_rows:
        log2n = args.log2_rows
    n = 1 << log2n
    tdt = torch.float32 if dt == "f32" else torch.float64
    code = cabi.LCC_F32 if dt == "f32" else cabi.LCC_F64
    alg = cabi.Algebra(*sig)
    nb = alg.num_blades
    flags = cabi.LCC_FLAG_STRICT_FP if args.strict else 0
    # a dedicated non-default stream: its handle is non-zero (NULL would mean "the library's own stream"),
    # kernels are launched on it and the CUDA events below are recorded on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0
    gen = torch.Generator(device=dev)
    gen.manual_seed(42 + rank)

    aos_layout = args.layout == "aos"
    if aos_layout and wl not in ("pga_sandwich", "cga_gp", "r3_gp", "cga_gp_outer"):
        raise SystemExit("--layout aos is available for pga_sandwich, cga_gp, cga_gp_outer and r3_gp")

    def soa(t):  # torch (nb, n) tensor == engine-native SoA view with ld = n
        if aos_layout:  # torch (n, nb) tensor == the reference's row-major layout (src/batch.rs:41-47)
            return cabi.view(t.data_ptr(), t.shape[0], t.stride(0), code, cabi.LCC_AOS)
        return cabi.view(t.data_ptr(), t.shape[1], t.stride(0), code, cabi.LCC_SOA)

    def soa_alloc(fill_random):  # (nb, n) blade-major view whose blade rows are n + --ld-pad elements apart
        t = torch.empty((nb, n + args.ld_pad), device=dev, dtype=tdt)
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

    motor = None

    def drain():  # workloads with asynchronous collectives override this
        pass

    if wl == "pga_sandwich":
        x = soa_alloc(False)
        pts = torch.randn((3, n), device=dev, dtype=tdt, generator=gen)
        x[7].fill_(1.0)
        x[14].copy_(-pts[0]); x[13].copy_(pts[1]); x[11].copy_(-pts[2])
        del pts
        if aos_layout:
            x = x.T.contiguous()
        out = torch.empty_like(x) if aos_layout else soa_alloc(False)
        motor = headline_motor()
        mptr = motor.ctypes.data_as(C.POINTER(C.c_double))
        vx, vo = soa(x), soa(out)

        def step():
            cabi.check(lib.lcc_batch_sandwich(alg.h, mptr, C.byref(vx), C.byref(vo), flags, sptr))
    elif wl == "pga_points_compact":
        x = torch.randn((n, 3), device=dev, dtype=tdt, generator=gen)
        out = torch.empty_like(x)
        motor = headline_motor()
        mptr = motor.ctypes.data_as(C.POINTER(C.c_double))

        def step():
            cabi.check(lib.lcc_pga_points_sandwich(alg.h, code, mptr, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), n, flags, sptr))
    elif wl in ("cga_gp", "r3_gp", "cga_gp_outer"):
        a = randn_batch()
        b = randn_batch()
        out = torch.empty_like(a) if aos_layout else soa_alloc(False)
        va, vb, vo = soa(a), soa(b), soa(out)
        if wl == "cga_gp_outer":
            out2 = torch.empty_like(a) if aos_layout else soa_alloc(False)
            vo2 = soa(out2)

            def step():
                cabi.check(lib.lcc_batch_gp_outer(alg.h, C.byref(va), C.byref(vb), C.byref(vo), C.byref(vo2), flags, sptr))
        else:
            def step():
                cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(va), C.byref(vb), C.byref(vo), flags, sptr))
    elif wl == "sta_fused_grade_inner_sum":
        a = torch.randn((nb, n), device=dev, dtype=tdt, generator=gen)
        b = torch.randn((nb, n), device=dev, dtype=tdt, generator=gen)
        # the per-GPU partial (16 values) is all-reduced asynchronously on NCCL's stream into a small ring of
        # result buffers, so the collective of step i overlaps the kernel of step i+1 instead of serialising with it
        ring = [torch.zeros(nb, device=dev, dtype=tdt) for _ in range(4)]
        pending = [None] * len(ring)
        turn = [0]
        sums = ring[0]
        out = None
        va, vb = soa(a), soa(b)

        def step():
            i = turn[0] % len(ring)
            turn[0] += 1
            if pending[i] is not None:
                pending[i].wait()  # the buffer's previous all-reduce (four steps ago) has completed
            cabi.check(lib.lcc_batch_product_sum(alg.h, cabi.LCC_OP_LC, C.byref(va), 1, C.byref(vb), C.c_void_p(ring[i].data_ptr()), flags, sptr))
            if world > 1:
                pending[i] = dist.all_reduce(ring[i], async_op=True)

        def drain():
            for h in pending:
                if h is not None:
                    h.wait()
    elif wl in ("cl8_fixed_gp", "cl8_fixed_gp_f64"):
        b = torch.randn((nb, n), device=dev, dtype=tdt, generator=gen)
        a = torch.randn((nb, 4), device=dev, dtype=tdt, generator=gen)  # the fixed operand: one-row SoA view, ld = 4
        out = torch.empty_like(b)
        va, vb, vo = cabi.view(a.data_ptr(), 1, 4, code, cabi.LCC_SOA), soa(b), soa(out)

        def step():
            cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(va), C.byref(vb), C.byref(vo), flags, sptr))
    elif wl == "sta_grade_inner_sum":
        a = torch.randn((nb, n), device=dev, dtype=tdt, generator=gen)
        b = torch.randn((nb, n), device=dev, dtype=tdt, generator=gen)
        g1 = torch.empty_like(a)
        out = torch.empty_like(a)
        sums = torch.zeros(nb, device=dev, dtype=tdt)
        va, vb, vg, vo = soa(a), soa(b), soa(g1), soa(out)

        def step():
            cabi.check(lib.lcc_batch_unary(alg.h, cabi.LCC_UN_GRADE, 1, C.byref(va), C.byref(vg), sptr))
            cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_LC, C.byref(vg), C.byref(vb), C.byref(vo), flags, sptr))
            cabi.check(lib.lcc_batch_sum(alg.h, C.byref(vo), C.c_void_p(sums.data_ptr()), sptr))
            if world > 1:
                dist.all_reduce(sums)  # the one collective of the path: 16 elements over NCCL
    else:
        raise SystemExit(f"unknown workload {wl}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    drain()
    barrier()
    lib.lcc_kernel_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        drain()  # every outstanding all-reduce is inside the timed region
        e1.record(stream)
        barrier()
    launches = int(lib.lcc_kernel_launch_count())
    elapsed_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * n / (ms_per_step / 1e3)

    # ---- parity spot check against the oracle on sampled rows (outside every timed region)
    parity = None
    if rank == 0 and not args.no_parity:
        import oracle

        o = oracle.OracleAlgebra(*sig)
        idx = torch.cat([torch.arange(0, 4096), torch.arange(n // 2, n // 2 + 4096), torch.arange(n - 4096, n)]).to(dev)
        tol = 1e-5 if dt == "f32" else 1e-12
        if wl == "pga_sandwich":
            xs = rows_of(x, idx)
            got = rows_of(out, idx)
            want = o.sandwich(motor, xs)
        elif wl == "pga_points_compact":
            xs, got = x[idx].cpu().numpy(), out[idx].cpu().numpy()
            want = oracle.pga_transform_points(motor, xs)
        elif wl == "sta_fused_grade_inner_sum":
            m = 1 << 20  # the fused kernel only returns the sums: re-run it on the first 2^20 rows and compare with the oracle
            sa, sb = cabi.view(a.data_ptr(), m, a.stride(0), code, cabi.LCC_SOA), cabi.view(b.data_ptr(), m, b.stride(0), code, cabi.LCC_SOA)
            chk = torch.zeros(nb, device=dev, dtype=tdt)
            cabi.check(lib.lcc_batch_product_sum(alg.h, cabi.LCC_OP_LC, C.byref(sa), 1, C.byref(sb), C.c_void_p(chk.data_ptr()), flags, sptr))
            torch.cuda.synchronize()
            xs, ys = a[:, :m].T.contiguous().cpu().numpy(), b[:, :m].T.contiguous().cpu().numpy()
            prod = o.left_contraction(o.grade(xs, 1), ys)
            want, mag = o.sum(prod)
            got = chk.cpu().numpy()
            want = want / np.maximum(mag, 1e-300) * np.abs(mag).max()  # compare on the scale of the summed magnitudes
            got = got / np.maximum(mag, 1e-300) * np.abs(mag).max()
        elif wl == "sta_grade_inner_sum":
            xs, ys = a[:, idx].T.contiguous().cpu().numpy(), b[:, idx].T.contiguous().cpu().numpy()
            got = out[:, idx].T.contiguous().cpu().numpy()
            want = o.left_contraction(o.grade(xs, 1), ys)
        elif wl in ("cl8_fixed_gp", "cl8_fixed_gp_f64"):
            idx = idx[::64]
            ys = b[:, idx].T.contiguous().cpu().numpy().astype(np.float64)
            got = out[:, idx].T.contiguous().cpu().numpy()
            want = o.geometric_product(a[:, 0].cpu().numpy().astype(np.float64)[None, :], ys)
        else:
            xs, ys = rows_of(a, idx), rows_of(b, idx)
            got = rows_of(out, idx)
            want = o.geometric_product(xs, ys)
        err = float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max() / np.abs(want).max())
        parity = {"rows_checked": int(idx.numel()), "max_rel_err": err, "tolerance": tol, "ok": bool(err <= tol)}

    # ---- end to end through the C-ABI host-buffer entry point (headline workload only)
    e2e = None
    if wl == "pga_sandwich" and not args.no_e2e:
        e2e = measure_e2e(args, lib, alg, code, motor, n, nb, world, rank, dev, dist if world > 1 else None, flags)
    elif wl == "pga_points_compact" and not args.no_e2e:
        e2e = measure_e2e(args, lib, alg, code, motor, n, nb, world, rank, dev, dist if world > 1 else None, flags, compact=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_kind = load_peaks()
    kernel_ms = ms_per_step  # one kernel per step for the headline; the step IS the kernel
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(wl)
    except Exception:
        pass
    cpu = None
    if not args.no_cpu and world == 1:  # CPU baseline on rank 0 at N = 1 only
        rows = cpu_sample_rows(wl)
        rate1, t1 = cpu_rate(wl, rows, 1)
        import oracle

        cores = oracle.max_threads()
        rate_all, _ = cpu_rate(wl, rows, cores) if cores > 1 else (rate1, t1)
        cpu = {"value": rate1, "unit": unit, "cores": 1, "kind": "port",
               "sample": f"{rows} rows of the same synthetic workload ({t1:.1f} s); the reference is single-threaded",
               "all_cores_value": rate_all, "all_cores": cores}
    line = {
        "metric": f"{wl} throughput", "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dt, "data": "synthetic",
        "config": {"workload": workload_name(wl), "rows_per_gpu": n, "layout": "dense (N,3) xyz rows" if wl == "pga_points_compact" else ("AoS (N, num_blades) row-major in HBM, TMA-staged tiles" if aos_layout else "SoA (blade-major) in HBM"),
                   "l2_policy": l2_policy(bytes_per * n), "strict_fp": bool(args.strict)},
        "roofline": roofline_block(wl, n, kernel_ms, bytes_per, peaks, peak_kind, traffic),
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(), "parity": parity,
    }
    emit_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def l2_policy(bytes_per_pass):
    gb = bytes_per_pass / 1e9
    if bytes_per_pass >= 8 * 126e6:
        return f"no flush needed: every step streams {gb:.1f} GB of distinct inputs and outputs through a 126 MB L2"
    return (f"NOT L2-cold: a step touches only {gb:.3f} GB, comparable to the 126 MB L2 -- this configuration is the reference's "
            "CPU-runnable case and is reported for completeness; use --log2-rows 26 for the HBM-bound figure")


def roofline_block(wl, n, kernel_ms, bytes_per, peaks, peak_kind, traffic):
    hbm = bytes_per * n / (kernel_ms / 1e3) / 1e9
    if wl == "cl8_fixed_gp_f64":  # FP64 tensor pipe: no measured figure in MEASURED_PEAKS.json, nominal B200 FP64 = 40 TFLOP/s
        tf = TENSOR_FLOPS_PER_UNIT[wl] * n / (kernel_ms / 1e3) / 1e12
        return {"bound": "tensor", "achieved": tf, "peak": 40.0, "unit": "TFLOP/s", "frac": tf / 40.0, "traffic": traffic,
                "peak_source": "nominal B200 FP64 tensor rate (40 TFLOP/s); MEASURED_PEAKS.json has no FP64 figure",
                "tensor_flops_per_unit": TENSOR_FLOPS_PER_UNIT[wl], "hbm_gbs": hbm, "hbm_frac": hbm / float(peaks["hbm_gbs"]),
                "algorithmic_bytes_per_unit": bytes_per}
    if wl in TENSOR_FLOPS_PER_UNIT:  # dense contraction: tensor-pipe bound; nominal dense tf32 peak is half the bf16 one
        tf = TENSOR_FLOPS_PER_UNIT[wl] * n / (kernel_ms / 1e3) / 1e12
        peak_tf32 = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
        sustained = float(peaks.get("bf16_tflops_sustained", 0.0)) / 2.0
        return {"bound": "tensor", "achieved": tf, "peak": peak_tf32, "unit": "TFLOP/s", "frac": tf / peak_tf32, "traffic": traffic,
                "peak_source": f"{peak_kind} bf16_tflops / 2 (burst figure; tf32 runs at half the bf16 rate; no direct tf32 measurement)",
                "peak_sustained": sustained or None, "frac_of_sustained": (tf / sustained) if sustained else None,
                "tensor_flops_per_unit": TENSOR_FLOPS_PER_UNIT[wl], "useful_tflops": tf / 3.0,
                "hbm_gbs": hbm, "hbm_frac": hbm / float(peaks["hbm_gbs"]), "algorithmic_bytes_per_unit": bytes_per}
    peak = float(peaks["hbm_gbs"])
    return {"bound": "hbm", "achieved": hbm, "peak": peak, "unit": "GB/s", "frac": hbm / peak, "traffic": traffic,
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
    """lcc_host_sandwich on pinned host buffers in the reference's (N, 16) row-major layout; with `compact`,
    lcc_host_pga_points_sandwich on pinned (N, 3) xyz buffers (the pga_points_compact workload)."""
    import psutil
    import torch

    numa_bound = bind_to_gpu_numa_node(dev.index or 0) if world > 1 else False

    from largecrimsoncanine_b200 import cabi

    es = 4 if code == cabi.LCC_F32 else 8
    if compact:
        return measure_e2e_compact(args, lib, alg, code, motor, n, es, world, rank, dev, dist, flags, numa_bound)
    n_e2e = n
    need = 2 * n_e2e * nb * es * max(world, 1)
    avail = psutil.virtual_memory().available
    while need > 0.5 * avail and n_e2e > (1 << 20):
        n_e2e //= 2
        need //= 2
    tdt = torch.float32 if es == 4 else torch.float64
    xh = torch.empty((n_e2e, nb), dtype=tdt, pin_memory=True)
    oh = torch.empty((n_e2e, nb), dtype=tdt, pin_memory=True)
    xh.zero_()
    g = torch.Generator().manual_seed(7 + rank)
    blk = 1 << 22
    for s in range(0, n_e2e, blk):  # fill in blocks: (N,3) normals embedded as PGA points
        e = min(n_e2e, s + blk)
        p = torch.randn((e - s, 3), dtype=tdt, generator=g)
        xh[s:e, 7] = 1.0
        xh[s:e, 14], xh[s:e, 13], xh[s:e, 11] = -p[:, 0], p[:, 1], -p[:, 2]
    mptr = motor.ctypes.data_as(C.POINTER(C.c_double))

    def call():
        cabi.check(lib.lcc_host_sandwich(alg.h, code, mptr, C.c_void_p(xh.data_ptr()), C.c_void_p(oh.data_ptr()), n_e2e, flags))

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

    o = oracle.OracleAlgebra(3, 0, 1)
    ok = bool(np.abs(oh[:2048].numpy() - o.sandwich(motor, xh[:2048].numpy())).max() <= 1e-4)
    return {"value": world * n_e2e * steps / dt_s, "unit": "points/s", "h2d_bytes_per_step": n_e2e * nb * es,
            "d2h_bytes_per_step": n_e2e * nb * es, "points_per_gpu": n_e2e, "steps": steps,
            "api": "lcc_host_sandwich (C ABI, pinned host buffers, chunked H2D/kernel/D2H pipeline)",
            "timer": "host wall clock around the blocking call, max over ranks", "checked_against_oracle": ok,
            "numa_bound": numa_bound}


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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
