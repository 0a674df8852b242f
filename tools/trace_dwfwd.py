"""In-kernel timeline (clock64) + graph-replay time of the fused depthwise-forward GEMM (tn_gemm_tc_dwfwd)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titanet_b200._ops import gemm_tc_raw
from titanet_b200._lib import LIB, call, ptr
B, T, C, Co, K = 64, 301, 256, 256, int(os.environ.get("K", 3))
P = float(os.environ.get("P", 0.1))
R = B * T
g = lambda *s: torch.randn(*s, device="cuda")
z, u, zo = g(R, C), torch.empty(R, C, device="cuda"), torch.empty(R, Co, device="cuda")
pw = g(Co, C) / 16; ws = torch.empty(4, Co, C, device="cuda"); call("tn_split_tf32", ptr(pw), ptr(ws), Co, C, 0)
dw_w, dw_b, pw_b = g(C, 1, K), g(C), g(Co)
sc, sh = torch.rand(C, device="cuda") + 0.5, g(C) * 0.1
seed = torch.tensor([5], dtype=torch.int64, device="cuda")
stats = torch.zeros(2 * Co, dtype=torch.float64, device="cuda")
def fused(): call("tn_gemm_tc_dwfwd", ptr(z), ptr(ws), ptr(dw_w), ptr(dw_b), ptr(sc), ptr(sh), 1, P, ptr(seed) if P > 0 else None, 3, ptr(pw_b), ptr(u), ptr(zo), None, None, B, T, C, Co, K, None)
def unfused():
    call("tn_dw_fwd", ptr(z), ptr(u), ptr(dw_w), ptr(dw_b), ptr(sc), ptr(sh), 1, P, ptr(seed) if P > 0 else None, 3, B, T, C, K)
    gemm_tc_raw(u, ws, pw_b, zo, stats, R, C, Co, 0, 3)
for name, f in (("fused", fused), ("dw_fwd + gemm", unfused)):
    f(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(gr):
            for _ in range(10): f()
    torch.cuda.synchronize(); gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) * 100:.1f} us per call (K={K}, p={P})")
tr = torch.zeros(1024, dtype=torch.int64, device="cuda")
LIB.call("tn_gemm_tc_set_trace", tr.data_ptr()); fused(); torch.cuda.synchronize(); LIB.call("tn_gemm_tc_set_trace", None)
t = tr.cpu().tolist(); t0 = t[0]; us = lambda v: (v - t0) / 1.9e3 if v else float("nan")
for kc in range(C // 32):
    print(f" chunk {kc}: tma_issue {us(t[1+kc]):6.2f}  full_seen {us(t[40+kc]):6.2f}  operand_done {us(t[20+kc]):6.2f}  ready_seen {us(t[60+kc]):6.2f}  mma_issued {us(t[80+kc]):6.2f}")
print(f" accum_seen {us(t[100]):6.2f}  epilogue_done {us(t[101]):6.2f}  exit {us(t[102]):6.2f}")
