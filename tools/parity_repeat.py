"""Repeat the S/17 batch-4 train-mode parity case N times in fresh model instances and print the embedding error vs the fp64
oracle each time (run-to-run spread = non-determinism of the forward pass) plus whether two runs agree bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import titanet_oracle as O
from cases import TRAIN_CASES, train_inputs
from test_gpu_model import build_model, rel

name = os.environ.get("CASE", "s17_ce_b4")
N = int(os.environ.get("N", "10"))
spec, loss, nc, B, T, scale, margin, full = TRAIN_CASES[name]
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
x, y = train_inputs(spec, nc, B, T)
sd64 = O.synth_state_dict(spec, loss, nc, dtype=torch.float64)
r64 = O.titanet_step(sd64, spec, x.double(), y, loss, scale=scale, margin=margin, input_grad=True)
ref_err = rel(g["emb"], r64[0])
errs, embs = [], []
for i in range(N):
    model = build_model(spec, loss, nc, scale, margin).train()
    emb, preds, lval = model(x.cuda(), speakers=y.cuda())
    errs.append(rel(emb, r64[0])); embs.append(emb.detach().cpu())
same = sum(int(torch.equal(embs[0], e)) for e in embs[1:])
print(f"{name} scheme3x={os.environ.get('TN_TC_3XTF32','0')}: fp32 reference err {ref_err:.3e}; ours min {min(errs):.3e} median {sorted(errs)[len(errs)//2]:.3e} "
      f"max {max(errs):.3e}; bit-identical to run 0: {same}/{N-1}")
print("  all:", " ".join(f"{e:.2e}" for e in errs))
