"""Launch-fused variants of the hot path, through the C ABI: train-mode BatchNorm folded by the last CTA of
its producer GEMM (tn_gemm_tc_bn / tn_conv_gemm_simt_bn), the one-pass BatchNorm backward (tn_bn_stats_bwd),
the per-step batched weight split (tn_split_tf32_batch) and the one-memset ZeroArena.  References: fp64 torch
restatements of conv -> nn.BatchNorm1d(train) -> ReLU (src/modules.py:119-134, src/models.py:452-455)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _bn(C, g):
    bn = torch.nn.BatchNorm1d(C)
    with torch.no_grad():
        bn.weight.copy_(0.5 + torch.rand(C, generator=g))
        bn.bias.copy_(0.3 * torch.randn(C, generator=g))
        bn.running_mean.copy_(torch.randn(C, generator=g))
        bn.running_var.copy_(0.5 + torch.rand(C, generator=g))
    return bn


@pytest.mark.parametrize("B,T,Ci,Co,K", [(8, 301, 256, 256, 1), (64, 301, 256, 256, 1), (3, 50, 80, 256, 3), (16, 1, 3072, 192, 1),
                                          (2, 300, 128, 1536, 1)])
def test_conv_bn_fused_matches_fp64(B, T, Ci, Co, K):
    """tcgen05 (1x1, many rows), CUDA-core (k=3 prolog shape) and split-K (decoder linear) producers."""
    from titanet_b200 import _ops as ops
    from titanet_b200 import _lib
    g = torch.Generator().manual_seed(B * T + Ci + Co + K)
    R = B * T
    x = torch.randn(R, Ci, generator=g)
    w = torch.randn(Co, Ci, K, generator=g) / math.sqrt(Ci * K)
    b = torch.randn(Co, generator=g)
    gy = torch.randn(R, Co, generator=g)
    bn_ref = _bn(Co, g).double().train()
    bn_gpu = _bn(Co, torch.Generator().manual_seed(0))
    bn_gpu.load_state_dict({k: v.float() for k, v in bn_ref.state_dict().items()})
    bn_gpu = bn_gpu.cuda().train()
    xg, wg, bg = (t.clone().cuda().requires_grad_(True) for t in (x, w, b))
    _lib.COUNTS.clear()
    z, sc, sh = ops.conv_gemm_bn(xg, wg, bg, bn_gpu, B, T)
    assert _lib.COUNTS.get("tn_bn_finalize", 0) == 0 and _lib.COUNTS.get("tn_colstats", 0) == 0   # no separate BatchNorm call
    # no ReLU here: with ~1e6 elements a few pre-activations sit within fp32 rounding of zero and an fp64 reference
    # takes the other branch (a legitimate, localised gradient difference); masks are covered by the other tests
    y = ops.Act.apply(z, sc, sh, None, False, 0.0, 0)
    (y * gy.cuda()).sum().backward()
    assert _lib.COUNTS.get("tn_bn_bwd_coef", 0) == 0 and _lib.COUNTS.get("tn_stats_bwd", 0) == 0
    # many rows and a 256-aligned channel count: the BatchNorm backward runs inside the data-gradient GEMM (tn_gemm_tc_bnbwd),
    # otherwise as one pass of its own (tn_bn_stats_bwd)
    fused = K == 1 and R >= 512 and Ci % 256 == 0 and Co % 32 == 0
    assert _lib.COUNTS.get("tn_bn_stats_bwd", 0) == (0 if fused else 1) and _lib.COUNTS.get("tn_gemm_tc_bnbwd", 0) == (1 if fused else 0)

    xr, wr, br = (t.double().clone().requires_grad_(True) for t in (x, w, b))
    xx = xr.view(B, T, Ci).permute(0, 2, 1)
    zr = F.conv1d(F.pad(xx, (K // 2, K // 2)), wr, br)
    yr = bn_ref(zr).permute(0, 2, 1).reshape(R, Co)
    (yr * gy.double()).sum().backward()
    assert rel(y, yr) < 2e-5
    assert rel(bn_gpu.running_mean, bn_ref.running_mean) < 1e-5 and rel(bn_gpu.running_var, bn_ref.running_var) < 1e-5
    assert int(bn_gpu.num_batches_tracked) == 1
    tol_w = 2e-3 if (K == 1 and R >= 256 and Ci % 32 == 0 and Co % 128 == 0) else 2e-4   # tensor-core wgrad: plain TF32 operands
    assert rel(xg.grad, xr.grad) < 2e-4
    assert rel(wg.grad, wr.grad) < tol_w
    assert rel(bn_gpu.weight.grad, bn_ref.weight.grad) < 2e-4
    assert rel(bn_gpu.bias.grad, bn_ref.bias.grad) < 2e-4
    # the conv bias in front of a train-mode BatchNorm has (mathematically) zero gradient
    assert float(bg.grad.abs().max()) < 1e-3 * float(gy.abs().sum(0).max())


def test_fused_fold_is_replayable():
    """The device-wide ticket resets itself: the same call twice gives the same fold (CUDA-graph replays)."""
    from titanet_b200 import _ops as ops
    g = torch.Generator().manual_seed(3)
    B, T, C = 8, 301, 256
    x = torch.randn(B * T, C, generator=g).cuda()
    w = (torch.randn(C, C, 1, generator=g) / 16).cuda()
    bn = _bn(C, g).cuda().train()
    outs = []
    for _ in range(3):
        z, sc, sh = ops.conv_gemm_bn(x, w, None, bn, B, T)
        outs.append((sc.clone(), sh.clone()))
    assert int(bn.num_batches_tracked) == 3
    for sc, sh in outs[1:]:
        assert rel(sc, outs[0][0]) < 1e-6 and rel(sh, outs[0][1]) < 1e-6


@pytest.mark.parametrize("B,T,C,Co,K,p", [(8, 301, 256, 256, 3, 0.0), (4, 77, 128, 256, 7, 0.0), (6, 301, 256, 256, 3, 0.1)])
def test_dwpw_bn_fused_matches_unfused(B, T, C, Co, K, p):
    """DwPwBN == DwPw + BNFold (our own unfused kernels), forward and every gradient."""
    from titanet_b200 import _ops as ops
    g = torch.Generator().manual_seed(B * T + C + K)
    R = B * T
    z = torch.randn(R, C, generator=g)
    sc, sh = 0.5 + torch.rand(C, generator=g), 0.3 * torch.randn(C, generator=g)
    dw_w, dw_b = torch.randn(C, 1, K, generator=g) / math.sqrt(K), torch.randn(C, generator=g)
    pw_w, pw_b = torch.randn(Co, C, 1, generator=g) / math.sqrt(C), torch.randn(Co, generator=g)
    gy = torch.randn(R, Co, generator=g).cuda()
    seed = torch.tensor([77], dtype=torch.int64, device="cuda")
    res = []
    for fused in (True, False):
        bn = _bn(Co, torch.Generator().manual_seed(9)).cuda().train()
        t = [x.clone().cuda().requires_grad_(True) for x in (z, sc, sh, dw_w, dw_b, pw_w, pw_b)]
        sd = seed if p > 0 else None
        if fused:
            zo, so, ho = ops.DwPwBN.apply(*t, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                                          bn.momentum, bn.eps, sd, True, p, 5, B, T)
        else:
            zo, st = ops.DwPw.apply(*t, sd, True, p, 5, B, T, True)
            so, ho = ops.bn_fold(st, bn, float(R))
        y = ops.Act.apply(zo, so, ho, None, True, 0.0, 0)
        (y * gy).sum().backward()
        res.append([y.detach()] + [x.grad for x in t] + [bn.weight.grad, bn.bias.grad, bn.running_mean.clone(), bn.running_var.clone()])
    names = ["y", "dz", "dscale", "dshift", "ddw", "ddb", "dpw", "dpb", "dgamma", "dbeta", "rmean", "rvar"]
    for n, a, b in zip(names, *res):
        if n in ("dpb", "ddb"):              # ~0 (biases in front of a train-mode BN): rounding noise, compare absolutely
            assert float((a - b).abs().max()) < 1e-2
        else:
            assert rel(a, b) < 2e-4, n


def test_model_step_uses_one_split_launch_and_arena_matches():
    """Whole model: one tn_split_tf32_batch per forward instead of per-GEMM splits; the ZeroArena step gives the
    same loss and gradients as the plain step."""
    import titanet_oracle as O
    from titanet_b200 import _lib, _ops as ops, losses, models, transforms
    from titanet_b200.engine import GraphedTrainStep
    spec = O.TitaNetSpec.named("s", 2)
    sd = O.synth_state_dict(spec, "ce", 251)
    wave, labels = O.synthetic_batch(4, seconds=1.0, seed=42)
    mel = transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=160, n_mels=80, specaugment_probability=0.0)
    results = []
    for mode in ("plain", "arena", "graph"):
        model = models.TitaNet.get_titanet(192, 80, 2, "s", loss_function=losses.CELoss(192, 251), dropout=0.0)
        model.load_state_dict(sd, strict=True)
        model = model.cuda().train()
        step = GraphedTrainStep(model, mel, 4, wave.shape[1], "cuda", use_graph=(mode == "graph"), use_arena=(mode != "plain"))
        _lib.COUNTS.clear()
        step.step(wave.cuda(), labels.cuda())
        loss = step.step(wave.cuda(), labels.cuda())
        torch.cuda.synchronize()
        if mode != "graph":
            assert _lib.COUNTS.get("tn_split_tf32", 0) == 0
            assert _lib.COUNTS.get("tn_split_tf32_batch", 0) == 2
            assert _lib.COUNTS.get("tn_bn_bwd_coef", 0) <= 2 * 1 and _lib.COUNTS.get("tn_stats_bwd", 0) <= 2 * 1   # decoder BN(3072) only
        if mode == "arena":
            assert step.arena.buf is not None
            base, size = step.arena.buf.data_ptr(), step.arena.buf.numel()
            inside = [0 <= p.grad.data_ptr() - base < size for p in model.parameters()]
            assert all(inside), f"{inside.count(False)} gradients live outside the arena"
        results.append((float(loss), {k: p.grad.detach().clone() for k, p in model.named_parameters()}))
    # the three modes run the same kernels; they differ only by the order of floating-point atomics, which train-mode
    # BatchNorm over a batch of 4 amplifies (the fp32 reference itself sits ~1e-2 from its fp64 run on such gradients)
    gmax = max(float(v.abs().max()) for v in results[0][1].values())
    for loss, grads in results[1:]:
        assert abs(loss - results[0][0]) < 1e-4 * abs(results[0][0])
        for k, gref in results[0][1].items():
            err = float((grads[k] - gref).abs().max()) / max(float(gref.abs().max()), 1e-2 * gmax)
            assert err < 5e-2, (k, err)


def test_optional_launch_fusions_match_default():
    """TN_FUSE_BLOCK_ENTRY / TN_FUSE_SE_MLP (off by default: measured slower) compute the same step (forward bit for bit:
    the fusions only touch the backward pass and the forward has no floating-point atomics)."""
    import titanet_oracle as O
    from titanet_b200 import _lib, _ops as ops, losses, models
    from cases import train_inputs
    spec = O.TitaNetSpec.named("s", 2)
    sd = O.synth_state_dict(spec, "ce", 251)
    x, y = train_inputs(spec, 251, 4, 101)
    res = []
    try:
        for fused in (False, True):
            ops.FUSE_BLOCK_ENTRY = ops.FUSE_SE_MLP = fused
            model = models.TitaNet.get_titanet(192, 80, 2, "s", loss_function=losses.CELoss(192, 251), dropout=0.0)
            model.load_state_dict(sd, strict=True)
            model = model.cuda().train()
            _lib.COUNTS.clear()
            emb, preds, loss = model(x.cuda(), speakers=y.cuda())
            loss.backward()
            assert (_lib.COUNTS.get("tn_tail_bwd1_mlp", 0) > 0) == fused
            res.append((float(loss), emb.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}))
    finally:
        ops.FUSE_BLOCK_ENTRY = ops.FUSE_SE_MLP = False
    assert abs(res[0][0] - res[1][0]) < 1e-4 * abs(res[0][0]) and rel(res[1][1], res[0][1]) < 1e-3   # atomics order x BatchNorm at batch 4
    gmax = max(float(v.abs().max()) for v in res[0][2].values())
    for k, gref in res[0][2].items():
        err = float((res[1][2][k] - gref).abs().max()) / max(float(gref.abs().max()), 1e-2 * gmax)
        assert err < 5e-2, (k, err)


@pytest.mark.parametrize("B,T,C,Co,K,lazy,p", [(8, 301, 256, 256, 3, True, 0.0), (8, 301, 256, 256, 3, True, 0.1), (5, 77, 128, 256, 7, True, 0.1),
                                                (64, 301, 256, 256, 3, False, 0.0), (3, 97, 64, 128, 5, True, 0.0), (2, 150, 256, 384, 1, True, 0.1)])
def test_depthwise_fused_into_gemm_operand(B, T, C, Co, K, lazy, p):
    """tn_gemm_tc_dwfwd (depthwise conv + BN/ReLU/dropout as the GEMM's operand producer) == tn_dw_fwd + tn_gemm_tc:
    same dropout masks (same counter-based hash), same u side output, same statistics; and == fp64 torch when p = 0."""
    from titanet_b200 import _ops as ops
    g = torch.Generator().manual_seed(B * T + C + K)
    R = B * T
    z = torch.randn(R, C, generator=g)
    sc, sh = 0.5 + torch.rand(C, generator=g), 0.3 * torch.randn(C, generator=g)
    dw_w, dw_b = torch.randn(C, 1, K, generator=g) / math.sqrt(K), torch.randn(C, generator=g)
    pw_w, pw_b = torch.randn(Co, C, 1, generator=g) / math.sqrt(C), torch.randn(Co, generator=g)
    seed = torch.tensor([91], dtype=torch.int64, device="cuda")
    outs = []
    default = ops.TC_FUSE_DWFWD
    try:
        for fused in (True, False):
            ops.TC_FUSE_DWFWD = fused
            stats = torch.empty(2 * Co, dtype=torch.float64, device="cuda")
            u, zo, _ = ops._dw_pw_forward(z.cuda(), sc.cuda() if lazy else None, sh.cuda() if lazy else None, dw_w.cuda(), dw_b.cuda(),
                                          pw_w.cuda(), pw_b.cuda(), seed if p > 0 else None, True, p, 7, B, T, stats, None)
            torch.cuda.synchronize()
            outs.append((u.clone(), zo.clone(), stats.clone()))
    finally:
        ops.TC_FUSE_DWFWD = default
    assert rel(outs[0][0], outs[1][0]) < 1e-6, "u side output"
    assert rel(outs[0][1], outs[1][1]) < 1e-5, "pointwise output"
    assert rel(outs[0][2][Co:], outs[1][2][Co:]) < 1e-5, "sum of squares"
    if p == 0.0:
        zz = z.double().view(B, T, C).permute(0, 2, 1)
        a = torch.relu(zz * sc.double().view(1, -1, 1) + sh.double().view(1, -1, 1)) if lazy else zz
        ur = F.conv1d(F.pad(a, (K // 2, K // 2)), dw_w.double(), dw_b.double(), groups=C)
        zr = F.conv1d(ur, pw_w.double(), pw_b.double()).permute(0, 2, 1).reshape(R, Co)
        assert rel(outs[0][0], ur.permute(0, 2, 1).reshape(R, C)) < 1e-5
        assert rel(outs[0][1], zr) < 1e-5


def test_prolog_conv_as_tensor_core_gemm_matches_fp64():
    """The K-tap dense conv of the prolog (ConvBlock1d(80, H, 3), src/models.py:370) with its taps unrolled into the reduction
    dimension (tn_im2col_nwc + tn_conv_weight_gemm) runs as one tcgen05 GEMM + BatchNorm fold; forward and the gradients of the
    weight, bias, gamma, beta against fp64 torch.  (An input that needs a gradient keeps the CUDA-core kernel.)"""
    from titanet_b200 import _lib, modules
    from titanet_b200.modules import Lazy
    B, T, Ci, Co, K = 64, 301, 80, 256, 3
    g = torch.Generator().manual_seed(11)
    x = 0.3 * torch.randn(B, Ci, T, generator=g)
    gy = torch.randn(B * T, Co, generator=g)
    conv = modules.Conv1dSamePadding(Ci, Co, K)
    bn_gpu = _bn(Co, g)
    ref_conv = torch.nn.Conv1d(Ci, Co, K, padding=K // 2).double()
    ref_conv.load_state_dict({k: v.double() for k, v in conv.state_dict().items()})
    bn_ref = _bn(Co, g).double().train()
    bn_ref.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in bn_gpu.state_dict().items()})
    conv, bn_gpu = conv.cuda(), bn_gpu.cuda().train()
    _lib.COUNTS.clear()
    z, sc, sh = conv._fwd_bn(Lazy.from_ncw(x.cuda()), bn_gpu)
    assert _lib.COUNTS.get("tn_im2col_nwc", 0) == 1 and _lib.COUNTS.get("tn_gemm_tc_bn", 0) == 1 and _lib.COUNTS.get("tn_conv_gemm_simt_bn", 0) == 0
    y = ops_act(z, sc, sh)
    (y * gy.cuda()).sum().backward()
    assert _lib.COUNTS.get("tn_wgrad_tc", 0) == 1
    yr = bn_ref(ref_conv(x.double())).permute(0, 2, 1).reshape(B * T, Co)
    (yr * gy.double()).sum().backward()
    assert rel(y, yr) < 1e-5
    assert rel(conv.weight.grad, ref_conv.weight.grad) < 2e-3                       # plain-TF32 weight-gradient GEMM
    assert rel(bn_gpu.weight.grad, bn_ref.weight.grad) < 1e-4 and rel(bn_gpu.bias.grad, bn_ref.bias.grad) < 1e-4
    assert rel(bn_gpu.running_var, bn_ref.running_var) < 1e-5
    gmax = float(ref_conv.weight.grad.abs().max())
    assert float(conv.bias.grad.abs().max()) < 1e-3 * gmax                           # mathematically zero in front of a train-mode BatchNorm


def ops_act(z, sc, sh):
    from titanet_b200 import _ops as ops
    return ops.Act.apply(z, sc, sh, None, False, 0.0, 0)
