"""Gradient error of the CUDA path vs the fp64 oracle for the tensor-core precision settings
(yard-stick: the fp32 oracle's own error vs fp64)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import titanet_oracle as O
from titanet_b200 import _ops as ops
from test_gpu_model import build_model

def rel(a, b, floor):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(floor))

def rel2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

B, blocks = int(os.environ.get("B", 16)), int(os.environ.get("BLOCKS", 17))
spec = O.TitaNetSpec.named("s", blocks)
g = torch.Generator().manual_seed(43)
x = 0.3 * torch.randn(B, 80, 301, generator=g); y = torch.randint(0, 251, (B,), generator=g)
r64 = O.titanet_step(O.synth_state_dict(spec, "ce", 251, dtype=torch.float64), spec, x.double(), y, "ce")
r32 = O.titanet_step(O.synth_state_dict(spec, "ce", 251), spec, x, y, "ce")
g64 = r64[3]
gmax = max(float(v.abs().max()) for v in g64.values())
def summarize(name, emb, loss, grads):
    worst = max(rel(grads[k], g64[k], 1e-3 * gmax) for k in g64)
    tot = (sum(float((grads[k].double().cpu() - g64[k]).norm() ** 2) for k in g64) / sum(float(g64[k].norm() ** 2) for k in g64)) ** 0.5
    print(f"{name:28s} emb {rel(emb, r64[0], 1e-30):.2e}  loss {abs(float(loss) - float(r64[2])) / float(r64[2]):.2e}  grad worst-tensor relmax {worst:.2e}  global rel-L2 {tot:.2e}")
summarize("fp32 oracle (CPU)", r32[0], r32[2], r32[3])
ONLY = os.environ.get("ONLY")          # e.g. ONLY="tc fwd3 dgrad3 wgrad1"; the split scheme itself follows TN_TC_3XTF32
for name, fwd, bwd, wg in [("simt fp32", 0, 0, 0), ("tc fwd3 dgrad3 wgradSIMT", 3, 3, 0), ("tc fwd3 dgrad3 wgrad1", 3, 3, 1), ("tc fwd3 dgrad1 wgrad1", 3, 1, 1), ("tc fwd1 dgrad1 wgrad1", 1, 1, 1)]:
    if ONLY and name != ONLY:
        continue
    ops.TC_ENABLED = fwd > 0
    ops.TC_FWD_NSPLIT, ops.TC_BWD_NSPLIT, ops.TC_WGRAD = max(fwd, 1), max(bwd, 1), wg > 0
    model = build_model(spec, "ce", 251).train()
    emb, preds, loss = model(x.cuda(), speakers=y.cuda())
    loss.backward()
    summarize(name, emb, loss, {k: p.grad for k, p in model.named_parameters()})
