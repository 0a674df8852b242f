"""Timeline (globaltimer) of the fused data-gradient kernel (BatchNorm backward as operand producer + pointwise dgrad GEMM +
depthwise / BN / ReLU / dropout backward epilogue) at the cfg-2 shape, plus graph-replayed time per launch with and without
the BatchNorm-backward producer."""
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titanet_b200 import _ops as ops
from titanet_b200._lib import LIB, TnBnBwd, call, ptr
B, T, C, K = int(os.environ.get("B", 64)), 301, int(os.environ.get("C", 256)), int(os.environ.get("K", 3))
P = float(os.environ.get("P", 0.1))
R = B * T
g = lambda *s: torch.randn(*s, device="cuda")
dz, zo, zprev, dzp, gout = g(R, C), g(R, C), g(R, C), torch.empty(R, C, device="cuda"), torch.empty(R, C, device="cuda")
pw = g(C, C) / math.sqrt(C); ws = torch.empty(4, C, C, device="cuda"); call("tn_split_tf32", ptr(pw), ptr(ws), C, C, 1)
dww = g(C, 1, K); ddw = torch.zeros(C, K, device="cuda"); acc = torch.zeros(4, C, device="cuda")
sc, sh = torch.rand(C, device="cuda") + 0.5, 0.1 * g(C)
seed = torch.tensor([1], dtype=torch.int64, device="cuda")
dsc, dsh, mean, invstd, gamma = 0.01 * g(C), 0.01 * g(C), 0.1 * g(C), torch.rand(C, device="cuda") + 0.5, torch.rand(C, device="cuda") + 0.5
dgam, dbet = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
bnb = TnBnBwd(ptr(zo), ptr(dsc), ptr(dsh), ptr(mean), ptr(invstd), ptr(gamma), float(R), ptr(gout), acc[3].data_ptr(), ptr(dgam), ptr(dbet))
def fused(): call("tn_gemm_tc_dwbwd_bn", ptr(dz), ptr(ws), ctypes.byref(bnb), ptr(zprev), ptr(dzp), ptr(dww), ptr(ddw), acc[0].data_ptr(), acc[1].data_ptr(),
                  acc[2].data_ptr(), ptr(sc), ptr(sh), 1, P, ptr(seed) if P > 0 else None, 3, B, T, C, C, K)
def plain(): call("tn_gemm_tc_dwbwd", ptr(dz), ptr(ws), ptr(zprev), ptr(dzp), ptr(dww), ptr(ddw), acc[0].data_ptr(), acc[1].data_ptr(),
                  acc[2].data_ptr(), ptr(sc), ptr(sh), 1, P, ptr(seed) if P > 0 else None, 3, B, T, C, C, K, 3)
def bnonly(): call("tn_bn_stats_bwd", ptr(dz), ptr(zo), ptr(dsc), ptr(dsh), ptr(mean), ptr(invstd), ptr(gamma), float(R), ptr(gout), acc[3].data_ptr(), ptr(dgam), ptr(dbet), R, C)
def timeit(name, f, reps=20):
    f(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(gr):
            for _ in range(reps): f()
    torch.cuda.synchronize(); gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    print(f"R={R} C={C} K={K} p={P} {name:40s} {e0.elapsed_time(e1) * 1e3 / reps:7.2f} us")
timeit("dgrad + dwbwd", plain)
timeit("bn_stats_bwd (separate)", bnonly)
timeit("bnbwd + dgrad + dwbwd (fused)", fused)
tr = torch.zeros(1024, dtype=torch.int64, device="cuda")
for _ in range(3): fused()
LIB.call("tn_gemm_tc_set_trace", tr.data_ptr()); fused(); torch.cuda.synchronize(); LIB.call("tn_gemm_tc_set_trace", None)
t = tr.cpu().tolist()
names = {0: "kernel entry", 1: "prologue done", 2: "transform loop done", 3: "accumulators complete", 7: "epilogue done", 8: "exit", 10: "warp0: producer loop done", 14: "warp0 done",
         20: "epilogue: z tiles in shared memory", 21: "epilogue: main loop done", 22: "epilogue: reductions out"}
for base, nm in ((0, "CTA 0"), (64, "CTA mid")):
    t0 = t[base]
    print(f"--- {nm}: us since kernel entry")
    for k in sorted(names, key=lambda k: t[base + k]):
        if t[base + k]: print(f"  {(t[base + k] - t0) / 1e3:7.2f}  {names[k]}")

kinds = ["TMA issued", "operands arrived (warp 2)", "producer done (warp 2)", "MMA: operands ready", "MMAs + commit issued"]
order = [0, 1, 2, 3, 4]
print("--- CTA 0 per K chunk, us since kernel entry: " + " | ".join(kinds))
for kc in range(min(16, C // 32)):
    print(f"  chunk {kc:2d}: " + " ".join(f"{(t[256 + 16 * k + kc] - t[0]) / 1e3:7.2f}" if t[256 + 16 * k + kc] else "      -" for k in order))
