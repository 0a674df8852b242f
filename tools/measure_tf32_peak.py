"""Dense TF32 peak of this GPU, measured the way MEASURED_PEAKS.json measures bf16 (torch.matmul 8192^3, 2*N^3 FLOP):
best of 10 (burst) and back to back for ~3 s (sustained).  Library GEMM (cuBLAS) as the yard-stick only.
Writes one JSON line; bench.py reads profiles/r02_tf32_peak.json when present."""
import json, sys, time
import torch

def main():
    torch.backends.cuda.matmul.allow_tf32 = True
    n = 8192
    a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2 * n ** 3 / (best * 1e-3) / 1e12
    t0 = time.time(); k = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(20):
            a @ b
        k += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sustained = 2 * n ** 3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
    # bf16 the same way, same box, for the ratio
    ah, bh = a.bfloat16(), b.bfloat16()
    for _ in range(3):
        ah @ bh
    torch.cuda.synchronize()
    bb = 1e9
    for _ in range(10):
        e0.record(); ah @ bh; e1.record(); torch.cuda.synchronize()
        bb = min(bb, e0.elapsed_time(e1))
    out = {"tf32_tflops": round(burst, 1), "tf32_tflops_sustained": round(sustained, 1),
           "bf16_tflops_same_box": round(2 * n ** 3 / (bb * 1e-3) / 1e12, 1),
           "how": "torch.matmul fp32 with allow_tf32, 8192^3: best of 10 (burst), back to back for 3 s (sustained)",
           "gpu": torch.cuda.get_device_name(0)}
    print(json.dumps(out))

if __name__ == "__main__":
    main()
