"""Graph-replay timing of the HBM-bound kernels at the cfg-2 shape (R=19264, C=256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titanet_b200._lib import call, ptr
B, T, C, K = 64, 301, 256, 3
R = B * T
g = lambda *s: torch.randn(*s, device="cuda")
z, du, dz, u, s_, out, dout = (g(R, C) for _ in range(7))
sc, sh = torch.rand(C, device="cuda") + 0.5, g(C) * 0.1
w, b = g(C, 1, K), g(C)
dw, db, dsc, dsh = torch.zeros(C, K, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
dst = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
seed = torch.tensor([123], dtype=torch.int64, device="cuda")
gate, dm = torch.rand(B, C, device="cuda"), g(B, C)
red = torch.zeros(4, C, device="cuda")
P = float(os.environ.get("P", 0.1))
fns = {
 "dw_fwd": lambda: call("tn_dw_fwd", ptr(z), ptr(u), ptr(w), ptr(b), ptr(sc), ptr(sh), 1, P, ptr(seed), 3, B, T, C, K),
 "dw_bwd": lambda: call("tn_dw_bwd", ptr(du), ptr(z), ptr(dz), ptr(w), ptr(dw), ptr(db), ptr(dsc), ptr(dsh), ptr(sc), ptr(sh), 1, P, ptr(seed), 3, B, T, C, K),
 "stats_bwd": lambda: call("tn_stats_bwd", ptr(du), ptr(z), ptr(dst), ptr(dz), ptr(db), R, C),
 "bn_stats_bwd": lambda: call("tn_bn_stats_bwd", ptr(du), ptr(z), ptr(dsc), ptr(dsh), ptr(sh), ptr(sc), ptr(sc), float(R), ptr(dz), ptr(db), ptr(dsc), ptr(dsh), R, C),
 "tail_bwd1": lambda: call("tn_tail_bwd1", ptr(dout), ptr(out), ptr(z), ptr(dm), ptr(sc), ptr(sh), P, 3, P, ptr(seed), B, T, C),
 "tail_fwd": lambda: call("tn_tail_fwd", ptr(z), ptr(s_), ptr(gate), ptr(out), ptr(sc), ptr(sh), P, 3, ptr(sc), ptr(sh), P, 4, ptr(seed), B, T, C),
 "tail_bwd2": lambda: call("tn_tail_bwd2", ptr(dout), ptr(out), ptr(z), ptr(s_), ptr(gate), ptr(dm), ptr(dz), ptr(du), red[0].data_ptr(), red[1].data_ptr(), red[2].data_ptr(), red[3].data_ptr(), ptr(sc), ptr(sh), P, 3, ptr(sc), ptr(sh), P, ptr(seed), B, T, C),
 "se_mean": lambda: call("tn_se_mean", ptr(z), ptr(gate), ptr(sc), ptr(sh), 1, P, ptr(seed), 3, B, T, C),
}
sel = sys.argv[1:] or list(fns)
for name in sel:
    f = fns[name]; f(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(gr):
            for _ in range(20): f()
    torch.cuda.synchronize(); gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{name:10s} P={P} RUN={os.environ.get('TN_DW_RUN','-')} RPB={os.environ.get('TN_EW_RPB','-')}: {e0.elapsed_time(e1) * 50:.2f} us")
