"""Print the in-kernel timeline (clock64) of the tensor-core GEMM for one shape."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titanet_b200._ops import gemm_tc_raw
from titanet_b200._lib import LIB, call, ptr
R, K, M, nsplit = (int(v) for v in (sys.argv[1:5] + ["19264", "256", "256", "3"][len(sys.argv) - 1:]))
x = torch.randn(R, K, device="cuda"); w = torch.randn(M, K, device="cuda") / math.sqrt(K)
z = torch.empty(R, M, device="cuda"); ws = torch.empty(4, M, K, device="cuda")
tr = torch.zeros(1024, dtype=torch.int64, device="cuda")
call("tn_split_tf32", ptr(w), ptr(ws), M, K, 0)
for _ in range(3): gemm_tc_raw(x, ws, None, z, None, R, K, M, 0, nsplit)
torch.cuda.synchronize()
LIB.load(); LIB.call("tn_gemm_tc_set_trace", tr.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
gemm_tc_raw(x, ws, None, z, None, R, K, M, 0, nsplit)
e1.record()
torch.cuda.synchronize(); LIB.call("tn_gemm_tc_set_trace", None)
print(f"event-timed launch: {e0.elapsed_time(e1) * 1e3:.1f} us")
t = tr.cpu().tolist(); nk = K // 32
for base, name in ((0, "CTA 0"), (128, "CTA mid")):
    t0 = t[base]
    us = lambda v: (v - t0) / 1.9e3 if v else float("nan")
    print(f"--- {name}: times in us since prologue end (clock64 / 1.9 GHz)")
    for kc in range(nk):
        print(f" chunk {kc}: tma_issue {us(t[base+1+kc]):6.2f}  full_seen {us(t[base+40+kc]):6.2f}  split_done {us(t[base+20+kc]):6.2f}  "
              f"ready_seen {us(t[base+60+kc]):6.2f}  mma_issued {us(t[base+80+kc]):6.2f}")
    print(f" kernel entry {us(t[base+110]):6.2f} (prologue = {-us(t[base+110]):.2f} us); globaltimer entry->exit {(t[base+112]-t[base+111])/1e3:.2f} us")
    print(f" accum_seen {us(t[base+100]):6.2f}  epilogue_done {us(t[base+101]):6.2f}  exit {us(t[base+102]):6.2f}")

import statistics
ent = [t[256 + 2 * i] for i in range(256) if t[256 + 2 * i]]; ext = [t[257 + 2 * i] for i in range(256) if t[257 + 2 * i]]
e0_ = min(ent)
print(f"CTAs traced {len(ent)}: entry spread {(max(ent)-e0_)/1e3:.2f} us (median {(statistics.median(ent)-e0_)/1e3:.2f}); last exit at {(max(ext)-e0_)/1e3:.2f} us; "
      f"lifetime median {statistics.median([b-a for a,b in zip(ent,ext)])/1e3:.2f} max {max(b-a for a,b in zip(ent,ext))/1e3:.2f} us")
