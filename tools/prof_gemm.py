"""Micro-harness: run one conv-GEMM shape a few times (for ncu / CUDA-event timing)."""
import argparse, math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titanet_b200._ops import gemm_tc_raw
from titanet_b200._lib import call, ptr

ap = argparse.ArgumentParser()
ap.add_argument("--R", type=int, default=19264); ap.add_argument("--K", type=int, default=256); ap.add_argument("--M", type=int, default=256)
ap.add_argument("--nsplit", type=int, default=3); ap.add_argument("--iters", type=int, default=20); ap.add_argument("--kind", default="tc"); ap.add_argument("--flags", type=int, default=0); ap.add_argument("--nostats", action="store_true")
a = ap.parse_args()
x = torch.randn(a.R, a.K, device="cuda"); w = torch.randn(a.M, a.K, device="cuda") / math.sqrt(a.K); b = torch.randn(a.M, device="cuda")
z = torch.empty(a.R, a.M, device="cuda"); st = torch.zeros(2 * a.M, device="cuda", dtype=torch.float64)
ws = torch.empty(4, a.M, a.K, device="cuda")
if a.nostats: st = None
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run():
    if a.kind == "tc":
        call("tn_split_tf32", ptr(w), ptr(ws), a.M, a.K, 0)
        gemm_tc_raw(x, ws, b, z, st, a.R, a.K, a.M, a.flags, a.nsplit)
    elif a.kind == "wgrad":
        call("tn_conv_wgrad_simt", ptr(z), ptr(x), ptr(ws), None, 1, a.R, a.K, a.M, 1)
for _ in range(3): run()
torch.cuda.synchronize()
ts = []
for _ in range(a.iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if a.kind == "tc": call("tn_split_tf32", ptr(w), ptr(ws), a.M, a.K, 0)
    e0.record()
    if a.kind == "tc": gemm_tc_raw(x, ws, b, z, st, a.R, a.K, a.M, a.flags, a.nsplit)
    else: run()
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
fl = 2.0 * a.R * a.K * a.M; by = 4.0 * (a.R * a.K + a.R * a.M + a.M * a.K)
print(f"{a.kind} R={a.R} K={a.K} M={a.M} nsplit={a.nsplit}: median {ts[len(ts)//2]:.1f} us  min {ts[0]:.1f} us  -> {fl/ts[len(ts)//2]/1e6:.1f} TFLOP/s, {by/ts[len(ts)//2]/1e3:.0f} GB/s algorithmic")
