"""Numerics of operand-split schemes for the fp32-equivalent tensor-core GEMM (CPU emulation, no GPU needed).

3xTF32 (what gemm_tc.cu issues today): x = xh + xl, w = wh + wl with xh, wh rounded to TF32 (10-bit mantissa),
    x.w ~= xh.wh + xl.wh + xh.wl                      -> 3 TF32 MMAs per product
TF32 + BF16 corrections (candidate): the two correction terms are ~2^-11 of the result, so their operands only need
    ~8 bits: xh.wh in TF32, [bf16(xl) | bf16(xh)] . [bf16(wh) ; bf16(wl)] as ONE bf16 MMA over a doubled K
    -> 1 TF32 MMA + 2 BF16 MMAs at twice the TF32 rate = 2 TF32-units instead of 3.
Prints the relative error of each scheme against fp64 for the TitaNet shapes (K = 256 / 512 / 1024).
"""
import torch


def tf32(x):
    """round-to-nearest-even to a 10-bit mantissa (cvt.rna.tf32.f32 rounds to nearest, ties away; the difference is immaterial here)"""
    i = x.contiguous().view(torch.int32)
    r = ((i >> 13) & 1) + 0x0FFF
    return ((i + r) & ~0x1FFF).view(torch.float32)


def bf16(x):
    return x.bfloat16().float()


def main():
    torch.manual_seed(0)
    for K in (256, 512, 1024):
        x = torch.randn(2048, K)
        x = torch.relu(x) * (torch.rand_like(x) > 0.1)           # post-ReLU, post-dropout activations
        w = torch.randn(256, K) / K ** 0.5
        ref = x.double() @ w.double().T
        xh, wh = tf32(x), tf32(w)
        xl, wl = tf32(x - xh), tf32(w - wh)
        mm = lambda a, b: (a.double() @ b.double().T)            # products exact, accumulation exact: isolates operand rounding
        schemes = {
            "fp32 (torch)": (x @ w.T).double(),
            "1xTF32": mm(xh, wh),
            "3xTF32": mm(xh, wh) + mm(xl, wh) + mm(xh, wl),
            "TF32 + BF16 corrections": mm(xh, wh) + mm(bf16(x - xh), bf16(wh)) + mm(bf16(xh), bf16(w - wh)),
        }
        scale = ref.abs().max()
        print(f"K = {K}")
        for name, got in schemes.items():
            err = (got - ref).abs()
            print(f"  {name:26s} max |err| / max |ref| = {float(err.max() / scale):.2e}   rms = {float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()):.2e}")


if __name__ == "__main__":
    main()
