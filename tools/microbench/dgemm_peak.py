"""cuBLAS DGEMM peak on the box (FP64 roofline denominator; MEASURED_PEAKS.json has no FP64 entry).

Run under gpurun:  python tools/microbench/dgemm_peak.py > gpurun_out/dgemm_peak.json
Burst = best of 10 single launches; sustained = back-to-back launches for ~3 s (power-capped clocks).
"""
import json
import subprocess
import sys
import time

import torch


def bench(m, n, k, reps=10, sustain_s=0.0):
    a = torch.randn(m, k, dtype=torch.float64, device="cuda")
    b = torch.randn(k, n, dtype=torch.float64, device="cuda")
    c = torch.empty(m, n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    out = {"m": m, "n": n, "k": k, "burst_tflops": 2.0 * m * n * k / best * 1e-12}
    if sustain_s > 0:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        iters = max(3, int(sustain_s / best))
        e0.record()
        for _ in range(iters):
            torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / iters
        out["sustained_tflops"] = 2.0 * m * n * k / t * 1e-12
        q = subprocess.run(
            ["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
             "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
        out["smi_after"] = q
    return out


if __name__ == "__main__":
    res = {"gpu": torch.cuda.get_device_name(0), "shapes": []}
    res["shapes"].append(bench(8192, 8192, 8192, sustain_s=3.0))
    for (m, n, k) in [(1024, 8192, 1024), (1024, 65536, 1024), (2048, 8192, 2048), (4096, 8192, 4096),
                      (256, 8192, 256), (4096, 65536, 4096)]:
        res["shapes"].append(bench(m, n, k))
    print(json.dumps(res, indent=1))
