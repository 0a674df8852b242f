"""Timeline (globaltimer, ns) of the pair GEMM with statistics + BatchNorm fold for one shape: two traced CTAs."""
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titanet_b200 import _ops as ops
from titanet_b200._lib import LIB, call, ptr
R, K, M = (int(v) for v in (sys.argv[1:4] + ["19264", "256", "256"][len(sys.argv) - 1:]))
x = torch.randn(R, K, device="cuda"); w = torch.randn(M, K, device="cuda") / math.sqrt(K); b = torch.randn(M, device="cuda")
z = torch.empty(R, M, device="cuda"); ws = torch.empty(4, M, K, device="cuda")
call("tn_split_tf32", ptr(w), ptr(ws), M, K, 0)
st = torch.empty(2 * M, dtype=torch.float64, device="cuda")
gamma, beta, rm, rv = torch.ones(M, device="cuda"), torch.zeros(M, device="cuda"), torch.zeros(M, device="cuda"), torch.ones(M, device="cuda")
nbt = torch.zeros((), dtype=torch.int64, device="cuda"); fold = torch.empty(4, M, device="cuda")
bn = ops.make_bn_fold(gamma, beta, rm, rv, nbt, 0.1, 1e-5, float(R), fold[0], fold[1], fold[2], fold[3])
sc, keep = ops.scratch(x)
def run(): call("tn_gemm_tc_bn", ptr(x), ptr(ws), ptr(b), ptr(z), ptr(st), ctypes.byref(bn), R, K, M, 0, 3, ctypes.byref(sc))
for _ in range(5): run()
torch.cuda.synchronize()
trs = [torch.zeros(1024, dtype=torch.int64, device="cuda") for _ in range(4)]
for _ in range(3): run()
for tr in trs:                # four launches back to back, one trace buffer each
    LIB.call("tn_gemm_tc_set_trace", tr.data_ptr())
    run()
torch.cuda.synchronize(); LIB.call("tn_gemm_tc_set_trace", None)
ts = [tr.cpu().tolist() for tr in trs]
for a, b in zip(ts[:-1], ts[1:]):
    print(f"launch-to-launch: entry(next, CTA 0) - entry(this, CTA 0) = {(b[0] - a[0]) / 1e3:.2f} us; exit(this, CTA 0 / mid) -> entry(next, CTA 0): "
          f"{(b[0] - a[8]) / 1e3:.2f} / {(b[0] - a[64 + 8]) / 1e3:.2f} us; entry skew CTA mid - CTA 0: {(a[64] - a[0]) / 1e3:.2f} us")
t = ts[-1]
names = {0: "kernel entry", 1: "prologue done (barriers, TMEM, cluster sync)", 2: "transform loop done", 3: "accumulators complete", 4: "stats pass done",
         5: "accumulator atomics issued", 6: "store pass starts", 7: "store pass done", 10: "warp0: producer loop done", 11: "warp0: barrier passed",
         12: "warp0: fence done", 13: "warp0: ticket returned", 14: "warp0: done (fold if last)", 8: "exit (after cluster sync)"}
for base, nm in ((0, "CTA 0"), (64, "CTA mid")):
    t0 = t[base]
    print(f"--- {nm}: us since kernel entry")
    for k in sorted(names, key=lambda k: t[base + k]):
        if t[base + k]: print(f"  {(t[base + k] - t0) / 1e3:7.2f}  {names[k]}")

# per-chunk events of CTA 0 (slots 256 + 16 kind + chunk)
kinds = ["TMA issued", "B arrived (warp 2)", "transform done (warp 2)", "MMA: operands of both CTAs ready", "MMAs + commit issued"]
order = [0, 1, 2, 3, 4]
t0 = t[0]
print("--- CTA 0 per K chunk, us since kernel entry: " + " | ".join(kinds))
for kc in range(min(16, K // 32)):
    print(f"  chunk {kc:2d}: " + " ".join(f"{(t[256 + 16 * k + kc] - t0) / 1e3:7.2f}" if t[256 + 16 * k + kc] else "      -" for k in order))
