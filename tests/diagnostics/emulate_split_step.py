"""CPU emulation of the split-GEMM arithmetic inside the oracle's training step (no GPU needed).

Every 1x1 convolution of the oracle (the ops the tcgen05 kernels run) is replaced by an autograd function whose forward
and data-gradient products use the emulated operand split (products and accumulation exact, so only operand rounding is
modelled) and whose weight gradient uses plain TF32 operands -- the configuration the CUDA path runs.  Prints the error of
each scheme against the fp64 oracle next to the fp32 oracle's own error, for the smoke() case and a deeper one.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")]
import torch
import torch.nn.functional as F
import titanet_oracle as O
from split_precision import bf16, tf32


def split_mm(a, b, scheme):
    """a [R, K] @ b [K, M] with emulated operand rounding."""
    if scheme == "fp32":
        return a @ b
    ah, bh = tf32(a), tf32(b.contiguous())
    d = lambda x, y: x.double() @ y.double()
    if scheme == "tf32":
        return d(ah, bh).float()
    if scheme == "3xtf32":
        return (d(ah, bh) + d(tf32(a - ah), bh) + d(ah, tf32(b - bh))).float()
    return (d(ah, bh) + d(bf16(a - ah), bf16(bh)) + d(bf16(ah), bf16(b - bh))).float()       # "tf32+bf16"


class Conv1x1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, scheme):                      # x [B, Ci, T], w [Co, Ci, 1]
        ctx.save_for_backward(x, w)
        ctx.scheme = scheme
        B, Ci, T = x.shape
        xr = x.permute(0, 2, 1).reshape(B * T, Ci)
        return split_mm(xr, w[:, :, 0].t(), scheme).reshape(B, T, -1).permute(0, 2, 1)

    @staticmethod
    def backward(ctx, dz):
        x, w = ctx.saved_tensors
        B, Ci, T = x.shape
        dzr = dz.permute(0, 2, 1).reshape(B * T, -1)
        xr = x.permute(0, 2, 1).reshape(B * T, Ci)
        dx = split_mm(dzr, w[:, :, 0], ctx.scheme).reshape(B, T, Ci).permute(0, 2, 1)
        dw = split_mm(dzr.t(), xr, "tf32" if ctx.scheme != "fp32" else "fp32")
        return dx, dw.unsqueeze(-1), None


def run(spec, B, T, scheme):
    orig = O.conv1d_same

    def patched(x, w, b, groups=1):
        if scheme and w.shape[-1] == 1 and groups == 1 and w.shape[0] % 128 == 0 and w.shape[1] % 32 == 0 and x.dtype == torch.float32:
            z = Conv1x1.apply(x, w, scheme)
            return z if b is None else z + b.view(1, -1, 1)
        return orig(x, w, b, groups)

    g = torch.Generator().manual_seed(43)
    x = 0.3 * torch.randn(B, 80, T, generator=g)
    y = torch.randint(0, 251, (B,), generator=g)
    O.conv1d_same = patched
    try:
        return O.titanet_step(O.synth_state_dict(spec, "ce", 251), spec, x, y, "ce"), (x, y)
    finally:
        O.conv1d_same = orig


def main():
    torch.set_num_threads(8)
    for blocks, B, T in ((2, 4, 101), (17, 4, 101)):
        spec = O.TitaNetSpec.named("s", blocks)
        r32, (x, y) = run(spec, B, T, None)
        r64 = O.titanet_step(O.synth_state_dict(spec, "ce", 251, dtype=torch.float64), spec, x.double(), y, "ce")
        k = "encoder.mega_blocks.0.sub_blocks.1.conv_block.0.conv.1.weight"
        print(f"TitaNet-S/{blocks}, B={B}, T={T}: error vs the fp64 oracle (embeddings rel-max | {k.split('.', 2)[2]} grad rel-max | all grads rel-L2)")
        for name, scheme in (("fp32 oracle", None), ("3xTF32", "3xtf32"), ("TF32 + BF16 corrections", "tf32+bf16"), ("plain TF32", "tf32")):
            r = r32 if scheme is None else run(spec, B, T, scheme)[0]
            e_emb = float((r[0].double() - r64[0]).abs().max() / r64[0].abs().max())
            e_k = float((r[3][k].double() - r64[3][k]).abs().max() / r64[3][k].abs().max())
            tot = (sum(float((r[3][n].double() - r64[3][n]).norm() ** 2) for n in r64[3]) / sum(float(r64[3][n].norm() ** 2) for n in r64[3])) ** 0.5
            print(f"  {name:26s} {e_emb:.2e} | {e_k:.2e} | {tot:.2e}")


if __name__ == "__main__":
    main()
