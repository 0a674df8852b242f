"""Per-launch GPU time of kernels replayed from a CUDA graph (no host gaps): same-kernel
back-to-back vs alternating with a small-shared-memory kernel."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titanet_b200._ops import gemm_tc_raw
from titanet_b200._lib import call, ptr
R, K, M = 19264, 256, 256
x = torch.randn(R, K, device="cuda"); w = torch.randn(M, K, device="cuda") / math.sqrt(K); b = torch.randn(M, device="cuda")
z = torch.empty(R, M, device="cuda"); ws = torch.empty(4, M, K, device="cuda"); st = torch.zeros(2 * M, device="cuda", dtype=torch.float64)
dst = torch.zeros(2 * M, dtype=torch.float64, device="cuda"); y = torch.empty(R, M, device="cuda"); dw = torch.zeros(M, K, device="cuda")
call("tn_split_tf32", ptr(w), ptr(ws), M, K, 0)
def gemm(n=3): gemm_tc_raw(x, ws, b, z, st, R, K, M, 0, n)
def small(): call("tn_stats_bwd", ptr(z), ptr(x), ptr(dst), ptr(y), None, R, M)
def tiny(): call("tn_split_tf32", ptr(w), ptr(ws), M, K, 0)
def wgrad(): call("tn_wgrad_tc", ptr(z), ptr(x), ptr(dw), R, K, M)
def timeit(name, fns, reps=20):
    for f in fns: f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for _ in range(reps):
                for f in fns: f()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) * 1e3 / reps:8.2f} us per iteration")
timeit("gemm3 x20", [gemm])
timeit("gemm1 x20", [lambda: gemm(1)])
timeit("stats_bwd x20", [small])
timeit("split x20", [tiny])
timeit("wgrad x20", [wgrad])
timeit("[stats_bwd, gemm3] x20", [small, gemm])
timeit("[split, gemm3] x20", [tiny, gemm])
timeit("[stats_bwd, wgrad] x20", [small, wgrad])
timeit("[gemm3, wgrad] x20", [gemm, wgrad])
