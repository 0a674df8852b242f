"""Per-tensor gradient error of a BIG_CASES entry (CUDA vs fp64 oracle, fp32 oracle as yard-stick): top offenders."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import titanet_oracle as O
from cases import BIG_CASES, big_inputs
from test_gpu_model import build_model, rel

from titanet_b200 import _ops as ops
if os.environ.get("SIMT_WGRAD") == "1":
    ops.TC_WGRAD = False            # exact fp32 weight gradients on the CUDA cores (A/B: is the plain-TF32 wgrad the offender?)
name = os.environ.get("CASE", "l5_arc_ragged_b4")
case = BIG_CASES[name]
spec, loss, nc = case["spec"], case["loss"], case["nc"]
x, y, frames = big_inputs(case)
kw = dict(scale=case["scale"], margin=case["margin"])
r64 = O.titanet_step(O.synth_state_dict(spec, loss, nc, dtype=torch.float64), spec, x.double(), y, loss, **kw)
r32 = O.titanet_step(O.synth_state_dict(spec, loss, nc), spec, x, y, loss, **kw)
model = build_model(spec, loss, nc, case["scale"], case["margin"]).train()
emb, preds, lval = model(x.cuda(), speakers=y.cuda())
lval.backward()
gmax = max(float(v.abs().max()) for v in r64[3].values())
rows = []
for k, p in model.named_parameters():
    rows.append((rel(p.grad, r64[3][k], floor=1e-3 * gmax), rel(r32[3][k], r64[3][k], floor=1e-3 * gmax), float(r64[3][k].abs().max()) / gmax, k))
rows.sort(reverse=True)
print(f"{name}: emb ours {rel(emb, r64[0]):.2e} fp32 oracle {rel(r32[0], r64[0]):.2e}; TN_TC_BWD_CORR={os.environ.get('TN_TC_BWD_CORR','1')} simt_wgrad={os.environ.get('SIMT_WGRAD','0')}")
for e, e32, mag, k in rows[:12]:
    print(f"  ours {e:.2e}  fp32-oracle {e32:.2e}  |g|max/gmax {mag:.1e}  {k}")
