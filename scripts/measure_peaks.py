#!/usr/bin/env python3
"""Micro-benchmark of the two tensor peaks MEASURED_PEAKS.json lacks: cuBLAS TF32 and
FP64 GEMM, 8192^3, best of 10 (burst) and back to back for 4 s (sustained), SM clock sampled under load; plus the bf16
figure and the copy bandwidth by the same method as the driver's file, for cross-checking this box against it.
Prints ONE JSON object; scripts/gpu.sh stage `peaks` stores it, and a reviewed copy lives at profiles/peaks_local.json,
which bench.py's roofline block reads."""
import json
import time

import torch

try:
    import pynvml

    pynvml.nvmlInit()
    H = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:
    pynvml = None

dev = torch.device("cuda", 0)
N = 8192


def clocks():
    if not pynvml:
        return None
    return pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM)


def gemm(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    a = torch.randn((N, N), device=dev, dtype=dtype)
    b = torch.randn((N, N), device=dev, dtype=dtype)
    c = torch.empty((N, N), device=dev, dtype=dtype)
    flop = 2.0 * N ** 3
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flop / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    # sustained: back to back for ~4 s
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    count, mhz = 0, []
    t_end = time.perf_counter() + 4.0
    e0.record()
    while time.perf_counter() < t_end:
        for _ in range(20):
            torch.matmul(a, b, out=c)
        count += 20
        torch.cuda.synchronize()
        mhz.append(clocks())
    e1.record()
    torch.cuda.synchronize()
    sustained = count * flop / (e0.elapsed_time(e1) * 1e-3) / 1e12
    mhz = [m for m in mhz if m]
    return {"burst_tflops": best, "sustained_tflops": sustained, "sm_mhz_under_load": (sorted(mhz)[len(mhz) // 2] if mhz else None)}


def copy_bw():
    a = torch.empty(1 << 30, device=dev, dtype=torch.bfloat16)
    b = torch.empty_like(a)
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2 * a.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


out = {
    "gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__, "n": N,
    "how": "torch.matmul 8192^3 (cuBLAS), 2*N^3 flop: best of 10 (burst) and back to back for 4 s (sustained); CUDA events",
    "tf32": gemm(torch.float32, True), "fp64": gemm(torch.float64, False), "fp32_no_tf32": gemm(torch.float32, False),
    "bf16": gemm(torch.bfloat16, False), "hbm_copy_gbs": copy_bw(),
    "sm_max_mhz": pynvml.nvmlDeviceGetMaxClockInfo(H, pynvml.NVML_CLOCK_SM) if pynvml else None,
}
print(json.dumps(out))
