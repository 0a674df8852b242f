"""Per-parameter gradient error of the smoke() model (TitaNet-S/2, batch 4, 1 s, CE, dropout 0) against the fp32 and fp64 CPU
oracles; env knobs (TN_TC_3XTF32, TN_FUSE_DWBWD, ...) select the CUDA configuration.  Prints the worst tensors."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import titanet_oracle as O
from titanet_b200 import losses, models, transforms

blocks = int(os.environ.get("BLOCKS", 2))
spec = O.TitaNetSpec.named("s", blocks)
sd = O.synth_state_dict(spec, "ce", 251)
model = models.TitaNet.get_titanet(192, 80, blocks, "s", loss_function=losses.CELoss(192, 251), dropout=0.0)
model.load_state_dict(sd, strict=True)
model = model.to("cuda:0").train()
wave, labels = O.synthetic_batch(4, seconds=1.0, seed=42)
x_ref = torch.cat([O.mel_spectrogram(w.view(1, -1)) for w in wave])
emb, preds, loss = model(x_ref.cuda(), speakers=labels.cuda())
loss.backward()
torch.cuda.synchronize()
r32 = O.titanet_step(sd, spec, x_ref, labels, "ce")
r64 = O.titanet_step(O.synth_state_dict(spec, "ce", 251, dtype=torch.float64), spec, x_ref.double(), labels, "ce")
gmax = max(float(v.abs().max()) for v in r64[3].values())
# rel-max per tensor with a floor of 1e-3 of the largest gradient entry of the model (conv biases in front of a BatchNorm have
# mathematically zero gradients: a plain relative error of those is meaningless)
rel = lambda a, b, floor=1e-30: float((a.detach().double().cpu() - b.double()).abs().max() / b.double().abs().max().clamp_min(floor))
grads = {k: p.grad for k, p in model.named_parameters()}
l2 = lambda gs: (sum(float((gs[k].detach().double().cpu() - r64[3][k]).norm() ** 2) for k in r64[3]) / sum(float(r64[3][k].norm() ** 2) for k in r64[3])) ** 0.5
rows = sorted(((rel(grads[k], r64[3][k], 1e-3 * gmax), rel(r32[3][k], r64[3][k], 1e-3 * gmax), k) for k in grads), reverse=True)
print(f"cfg 3xtf32={os.environ.get('TN_TC_3XTF32', '0')} fuse_dwbwd={os.environ.get('TN_FUSE_DWBWD', '1')}: emb {rel(emb, r64[0]):.2e} "
      f"(fp32 oracle {rel(r32[0], r64[0]):.2e})  all grads rel-L2 {l2(grads):.2e} (fp32 oracle {l2(r32[3]):.2e})")
for e, e32, k in rows[:5]:
    print(f"   {e:.2e} (fp32 oracle {e32:.2e})  {k}")
k = "encoder.mega_blocks.0.sub_blocks.1.conv_block.0.conv.1.weight"          # the tensor smoke() looks at
print(f"   smoke tensor: max|g| = {float(r64[3][k].abs().max()) / gmax:.2e} of the model's largest gradient entry; rel-max (own scale) vs fp64: "
      f"CUDA {rel(grads[k], r64[3][k]):.2e}, fp32 oracle {rel(r32[3][k], r64[3][k]):.2e}; CUDA vs fp32 oracle {rel(grads[k], r32[3][k]):.2e}; "
      f"rel-L2 vs fp64: CUDA {float((grads[k].double().cpu() - r64[3][k]).norm() / r64[3][k].norm()):.2e}, fp32 oracle {float((r32[3][k].double() - r64[3][k]).norm() / r64[3][k].norm()):.2e}")
