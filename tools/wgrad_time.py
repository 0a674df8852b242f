import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titanet_b200._lib import call, ptr
R, Ci, Co = 19264, 256, 256
dz, u = torch.randn(R, Co, device="cuda"), torch.randn(R, Ci, device="cuda"); dw = torch.zeros(Co, Ci, device="cuda")
f = lambda: call("tn_wgrad_tc", ptr(dz), ptr(u), ptr(dw), R, Ci, Co)
f(); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
with torch.cuda.stream(st):
    with torch.cuda.graph(gr):
        for _ in range(20): f()
torch.cuda.synchronize(); gr.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
print(f"wgrad MT={os.environ.get('TN_WG_MT','-')} NB={os.environ.get('TN_WG_NB','-')}: {e0.elapsed_time(e1) * 50:.2f} us")
