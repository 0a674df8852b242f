"""Eager launches of the three tensor-core kernels at the cfg-2 shape for `ncu --set full` (3 launches each after warm-up):
forward pair GEMM with statistics + BatchNorm fold, fused BatchNorm-backward + dgrad + depthwise-backward, weight gradient."""
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titanet_b200 import _ops as ops
from titanet_b200._lib import TnBnBwd, call, ptr
B, T, C, K = 64, 301, 256, 3
R = B * T
g = lambda *s: torch.randn(*s, device="cuda")
x, z = g(R, C), torch.empty(R, C, device="cuda")
w = g(C, C) / math.sqrt(C); b = g(C)
ws = torch.empty(4, C, C, device="cuda"); call("tn_split_tf32", ptr(w), ptr(ws), C, C, 0)
wst = torch.empty(4, C, C, device="cuda"); call("tn_split_tf32", ptr(w), ptr(wst), C, C, 1)
st = torch.empty(2 * C, dtype=torch.float64, device="cuda"); fold = torch.empty(4, C, device="cuda")
bnp = [torch.ones(C, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")]
bn = ops.make_bn_fold(bnp[0], bnp[1], bnp[2], bnp[3], bnp[4], 0.1, 1e-5, float(R), fold[0], fold[1], fold[2], fold[3])
sc, keep = ops.scratch(x)
dz, zo, zprev, dzp, gout = g(R, C), g(R, C), g(R, C), torch.empty(R, C, device="cuda"), torch.empty(R, C, device="cuda")
dww = g(C, 1, K); ddw = torch.zeros(C, K, device="cuda"); acc = torch.zeros(4, C, device="cuda")
scl, shf = torch.rand(C, device="cuda") + 0.5, 0.1 * g(C)
seed = torch.tensor([1], dtype=torch.int64, device="cuda")
v = [0.01 * g(C), 0.01 * g(C), 0.1 * g(C), torch.rand(C, device="cuda") + 0.5, torch.rand(C, device="cuda") + 0.5, torch.empty(C, device="cuda"), torch.empty(C, device="cuda")]
bnb = TnBnBwd(ptr(zo), ptr(v[0]), ptr(v[1]), ptr(v[2]), ptr(v[3]), ptr(v[4]), float(R), ptr(gout), acc[3].data_ptr(), ptr(v[5]), ptr(v[6]))
dW = torch.zeros(C, C, device="cuda")
def fwd(): call("tn_gemm_tc_bn", ptr(x), ptr(ws), ptr(b), ptr(z), ptr(st), ctypes.byref(bn), R, C, C, 0, 3, ctypes.byref(sc))
def bwd(): call("tn_gemm_tc_dwbwd_bn", ptr(dz), ptr(wst), ctypes.byref(bnb), ptr(zprev), ptr(dzp), ptr(dww), ptr(ddw), acc[0].data_ptr(), acc[1].data_ptr(),
                acc[2].data_ptr(), ptr(scl), ptr(shf), 1, 0.1, ptr(seed), 3, B, T, C, C, K)
def wg(): call("tn_wgrad_tc", ptr(gout), ptr(x), ptr(dW), R, C, C)
for f in (fwd, bwd, wg):
    for _ in range(4): f()
torch.cuda.synchronize()
