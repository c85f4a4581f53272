#!/usr/bin/env python
"""Measure the TF32 tensor-core GEMM peak of this GPU the way MEASURED_PEAKS.json measures bf16 (torch.matmul 8192^3,
best of 10 = burst; back to back for 4 s = sustained) and write profiles/tf32_peak.json.  bench.py uses the sustained
figure as the roofline denominator of the TF32 scoring kernels when the file exists (otherwise bf16_sustained / 2).

    gpurun -- 'python benchmarks/measure_tf32_peak.py'   # then copy gpurun_out/tf32_peak.json to profiles/
"""
import json
import os
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    torch.backends.cuda.matmul.allow_tf32 = True
    n = 8192
    a = torch.randn((n, n), device="cuda", dtype=torch.float32)
    b = torch.randn((n, n), device="cuda", dtype=torch.float32)
    flop = 2.0 * n ** 3
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flop / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, reps = time.time(), 0
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            a @ b
        reps += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = reps * flop / (e0.elapsed_time(e1) * 1e-3) / 1e12
    out = {"tf32_tflops": best, "tf32_tflops_sustained": sustained, "gpu_name": torch.cuda.get_device_name(0),
           "how": "torch.matmul fp32 with allow_tf32, 8192^3: best of 10 (burst), back to back for 4 s (sustained)",
           "torch": torch.__version__}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tf32_peak.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
