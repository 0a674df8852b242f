"""Error of tn_gemm_tc (nsplit = 3) against fp64 for the shapes of the smoke() model (R = 404) and of the benchmark; the split
scheme follows TN_TC_FWD_CORR (0: 3xTF32, 1: TF32 + bf16 correction, 2 = default: TF32 + scaled-fp16 correction).  Prints rel-max and rms-relative error per shape."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from titanet_b200._ops import gemm_tc_raw
from titanet_b200._lib import call, ptr

def run(R, Kd, M, transpose, relu, flags=0):
    g = torch.Generator().manual_seed(R + Kd + M)
    x = torch.randn(R, Kd, generator=g)
    if relu:
        x = torch.relu(x)
    w = torch.randn(*( (Kd, M) if transpose else (M, Kd) ), generator=g) / math.sqrt(Kd)
    ref = x.double() @ (w.double() if transpose else w.double().t())
    ws = torch.empty(4, M, Kd, device="cuda")
    call("tn_split_tf32", ptr(w.cuda()), ptr(ws), M, Kd, int(transpose))
    z = torch.empty(R, M, device="cuda")
    gemm_tc_raw(x.cuda(), ws, None, z, None, R, Kd, M, flags, 3)
    torch.cuda.synchronize()
    err = (z.double().cpu() - ref)
    print(f"R={R:6d} Kd={Kd:5d} M={M:5d} T={int(transpose)} relu={int(relu)}: relmax {float(err.abs().max() / ref.abs().max()):.2e}  rms {float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()):.2e}")

print("scheme: TN_TC_FWD_CORR =", os.environ.get("TN_TC_FWD_CORR", "2 (default)"))
for R in (404, 19264):
    for Kd, M in ((256, 256), (256, 1536), (1536, 256), (1536, 128), (128, 1536)):
        for tr in (False, True):
            run(R, Kd, M, tr, True)
