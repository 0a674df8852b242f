"""Timeline (globaltimer, ns) of the weight-gradient kernel wgrad_tc_kernel at the cfg-2 shape: first and last split's CTA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titanet_b200._lib import LIB, call, ptr
R, Ci, Co = (int(v) for v in (sys.argv[1:4] + ["19264", "256", "256"][len(sys.argv) - 1:]))
dz, u = torch.randn(R, Co, device="cuda"), torch.randn(R, Ci, device="cuda"); dw = torch.zeros(Co, Ci, device="cuda")
big = torch.empty(64 << 20, device="cuda")             # 256 MB: written between launches so that the operands come from HBM
f = lambda: call("tn_wgrad_tc", ptr(dz), ptr(u), ptr(dw), R, Ci, Co)
for _ in range(3): f()
torch.cuda.synchronize()
names = {0: "kernel entry", 1: "prologue done", 2: "producer: last chunk issued", 3: "MMA: first chunk issued", 4: "MMA: all issued",
         5: "accumulators complete", 6: "tile staged in shared memory", 7: "reductions issued and read", 8: "exit"}
for cold in (0, 1):
    tr = torch.zeros(64, dtype=torch.int64, device="cuda")
    if cold: big.fill_(1.0)
    LIB.call("tn_gemm_tc_set_trace", tr.data_ptr()); f(); torch.cuda.synchronize(); LIB.call("tn_gemm_tc_set_trace", None)
    t = tr.cpu().tolist()
    for base, nm in ((0, "first split"), (16, "last split")):
        print(f"--- {'operands from HBM' if cold else 'operands in L2'}, {nm}: us since kernel entry of the first split")
        for k in sorted(names, key=lambda k: t[base + k]):
            if t[base + k]: print(f"  {(t[base + k] - t[0]) / 1e3:7.2f}  {names[k]}")
