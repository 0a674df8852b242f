"""Graph-replayed time per launch of the forward pair GEMM (R=19264, 256 -> 256 by default) with / without statistics and the
BatchNorm fold; the split scheme follows TN_TC_FWD_CORR (0: 3xTF32, 1: TF32 + bf16 correction, 2 = default: TF32 + scaled-fp16 correction)."""
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titanet_b200 import _ops as ops
from titanet_b200._lib import LIB, call, ptr
R, K, M = (int(v) for v in (sys.argv[1:4] + ["19264", "256", "256"][len(sys.argv) - 1:]))
dev = "cuda"
x = torch.randn(R, K, device=dev); w = torch.randn(M, K, device=dev) / math.sqrt(K); b = torch.randn(M, device=dev)
z = torch.empty(R, M, device=dev); ws = torch.empty(4, M, K, device=dev)
call("tn_split_tf32", ptr(w), ptr(ws), M, K, 0)
st = torch.empty(2 * M, dtype=torch.float64, device=dev)
gamma, beta, rm, rv = torch.ones(M, device=dev), torch.zeros(M, device=dev), torch.zeros(M, device=dev), torch.ones(M, device=dev)
nbt = torch.zeros((), dtype=torch.int64, device=dev); fold = torch.empty(4, M, device=dev)
bn = ops.make_bn_fold(gamma, beta, rm, rv, nbt, 0.1, 1e-5, float(R), fold[0], fold[1], fold[2], fold[3])
sc, keep = ops.scratch(x)
def plain(): call("tn_gemm_tc", ptr(x), ptr(ws), ptr(b), ptr(z), None, R, K, M, 0, 3, None)
def grad(): call("tn_gemm_tc", ptr(x), ptr(ws), ptr(b), ptr(z), None, R, K, M, 8, 3, None)
def stats(): call("tn_gemm_tc", ptr(x), ptr(ws), ptr(b), ptr(z), ptr(st), R, K, M, 0, 3, ctypes.byref(sc))
def bnf(): call("tn_gemm_tc_bn", ptr(x), ptr(ws), ptr(b), ptr(z), ptr(st), ctypes.byref(bn), R, K, M, 0, 3, ctypes.byref(sc))
def timeit(name, f, reps=20):
    f(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for _ in range(reps): f()
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    print(f"R={R} K={K} M={M} fwd_corr={os.environ.get('TN_TC_FWD_CORR', '2')} nacc_cap={os.environ.get('TN_TC_NACC', '8')} {name:28s} {best:7.2f} us")
timeit("plain (no stats)", plain)
timeit("gradient flavour", grad)
timeit("+ statistics", stats)
timeit("+ statistics + BN fold", bnf)
