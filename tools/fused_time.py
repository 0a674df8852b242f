"""Graph-replay timing + in-kernel timeline of the fused dgrad + depthwise-backward kernel."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titanet_b200._lib import LIB, call, ptr
B, T, C, Co, K = 64, 301, 256, 256, 3
R = B * T
g = lambda *s: torch.randn(*s, device="cuda")
dz, z, dzp = g(R, Co), g(R, C), torch.empty(R, C, device="cuda")
pw = g(Co, C) / 16; ws = torch.empty(4, C, Co, device="cuda"); call("tn_split_tf32", ptr(pw), ptr(ws), C, Co, 1)
dw_w = g(C, 1, K); ddw = torch.zeros(C, K, device="cuda"); ddb, dsc, dsh = (torch.zeros(C, device="cuda") for _ in range(3))
sc, sh = torch.rand(C, device="cuda") + 0.5, g(C) * 0.1
seed = torch.tensor([5], dtype=torch.int64, device="cuda")
def run(p, lazy):
    call("tn_gemm_tc_dwbwd", ptr(dz), ptr(ws), ptr(z), ptr(dzp), ptr(dw_w), ptr(ddw), ptr(ddb), ptr(dsc) if lazy else None, ptr(dsh) if lazy else None,
         ptr(sc) if lazy else None, ptr(sh) if lazy else None, 1, p, ptr(seed) if p > 0 else None, 3, B, T, Co, C, K, 3)
tr = torch.zeros(1024, dtype=torch.int64, device="cuda")
for p, lazy in ((0.1, True), (0.0, True), (0.0, False)):
    run(p, lazy); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(gr):
            for _ in range(10): run(p, lazy)
    torch.cuda.synchronize(); gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    LIB.call("tn_gemm_tc_set_trace", tr.data_ptr()); tr.zero_(); run(p, lazy); torch.cuda.synchronize(); LIB.call("tn_gemm_tc_set_trace", None)
    t = tr.cpu().tolist(); us = lambda v: (v - t[0]) / 1.9e3
    print(f"p={p} lazy={lazy}: {e0.elapsed_time(e1) * 100:.1f} us/launch; CTA0: accum_seen {us(t[100]):.1f} us, epilogue_done {us(t[101]):.1f} us, "
          f"kernel_end {us(t[102]):.1f}; warp2 mt0 [z ready {us(t[113]):.1f}, primed {us(t[114]):.1f}, loop done {us(t[115]):.1f}, atomics issued {us(t[116]):.1f}] "
          f"mt1 [{us(t[117]):.1f}, {us(t[118]):.1f}, {us(t[119]):.1f}, {us(t[120]):.1f}]; globaltimer CTA span {(t[112] - t[111]) / 1e3:.1f} us")
