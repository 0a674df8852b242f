"""tcgen05 (tensor-core) GEMM against an fp64 CPU product, through the C ABI."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def run_tc(x, w, bias, nsplit, transpose=False, want_stats=True, flags=0, z0=None):
    import ctypes
    from titanet_b200._lib import LIB, call, ptr
    from titanet_b200._ops import scratch
    R, Kd = x.shape
    M = w.shape[1] if transpose else w.shape[0]
    ws = torch.empty(4, M, Kd, device="cuda")
    call("tn_split_tf32", ptr(w), ptr(ws), M, Kd, int(transpose))
    z = torch.empty(R, M, device="cuda") if z0 is None else z0
    stats = torch.full((2 * M,), float("nan"), device="cuda", dtype=torch.float64) if want_stats else None   # written, not accumulated
    sc, keep = scratch(x) if want_stats else (None, None)
    call("tn_gemm_tc", ptr(x), ptr(ws), ptr(bias), ptr(z), ptr(stats), R, Kd, M, flags, nsplit, ctypes.byref(sc) if sc is not None else None)
    torch.cuda.synchronize()
    return z, stats, ws


@pytest.mark.parametrize("R,Kd,M", [(19264, 256, 256), (301, 256, 256), (1000, 128, 1536), (777, 1536, 128), (4096, 512, 512),
                                     (2000, 256, 1536), (515, 64, 384)])
def test_gemm_tc_3xtf32_matches_fp64(R, Kd, M):
    g = torch.Generator().manual_seed(R + Kd + M)
    x = torch.randn(R, Kd, generator=g)
    w = torch.randn(M, Kd, generator=g) / math.sqrt(Kd)
    b = torch.randn(M, generator=g)
    ref = x.double() @ w.double().t() + b.double()
    z, stats, ws = run_tc(x.cuda(), w.cuda(), b.cuda(), 3)
    # weight split: ws[0] = tf32(w); ws[1] = tf32(w - hi) (3xTF32, forward GEMMs); ws[2] = per row and 32-wide K chunk, 64 bf16
    # = [bf16(hi) x32 | bf16(w - hi) x32] (the weight side of the bf16 correction MMA of the gradient GEMMs)
    hi = ws[0].cpu()
    assert rel(hi, w) < 6e-4
    assert rel(ws[0] + ws[1], w) < 1e-6
    corr = ws[2].cpu().view(torch.bfloat16).view(M, Kd // 32, 64).float()
    assert torch.equal(corr[:, :, :32].reshape(M, Kd), hi.bfloat16().float())
    assert torch.equal(corr[:, :, 32:].reshape(M, Kd), (w - hi).bfloat16().float())
    assert rel(z, ref) < 1e-5, "the split GEMM must be fp32-equivalent (torch fp32 itself sits at ~5e-7)"
    # gradient flavour (TF32 + one bf16 correction MMA), same product
    zg, _, _ = run_tc(x.cuda(), w.cuda(), b.cuda(), 3, want_stats=False, flags=8)
    assert rel(zg, ref) < 2e-5
    # bit-for-bit reproducible, statistics included (no floating-point atomics in the forward path)
    z2, stats2, _ = run_tc(x.cuda(), w.cuda(), b.cuda(), 3)
    assert torch.equal(z, z2) and torch.equal(stats, stats2)
    assert rel(stats[:M], ref.sum(0)) < 1e-4 * max(1.0, float(ref.abs().sum(0).max() / ref.sum(0).abs().max()))
    assert rel(stats[M:], (ref ** 2).sum(0)) < 1e-5


def test_gemm_tc_plain_tf32_and_transpose_and_flags():
    R, Kd, M = 3000, 256, 256
    g = torch.Generator().manual_seed(5)
    x = torch.randn(R, Kd, generator=g)
    w = torch.randn(Kd, M, generator=g) / math.sqrt(Kd)          # stored [Kd, M]: the dgrad layout
    ref = x.double() @ w.double()
    z1, _, _ = run_tc(x.cuda(), w.cuda(), None, 1, transpose=True, want_stats=False)
    assert 1e-5 < rel(z1, ref) < 3e-3, "plain TF32: ~1e-3 accurate"
    z3, _, _ = run_tc(x.cuda(), w.cuda(), None, 3, transpose=True, want_stats=False)
    assert rel(z3, ref) < 1e-5
    # accumulate + tanh epilogues
    z0 = torch.randn(R, M, generator=g)
    za, _, _ = run_tc(x.cuda(), w.cuda(), None, 3, transpose=True, want_stats=False, flags=2, z0=z0.cuda().clone())
    assert rel(za, ref + z0.double()) < 1e-5
    zt, _, _ = run_tc(x.cuda(), w.cuda(), None, 3, transpose=True, want_stats=False, flags=1)
    assert rel(zt, torch.tanh(ref)) < 1e-5


def test_conv_gemm_op_uses_tensor_cores_and_matches_simt():
    from titanet_b200 import _ops as ops
    from titanet_b200 import _lib
    B, T, Ci, Co = 8, 301, 256, 256
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B * T, Ci, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(Co, Ci, 1, generator=g) / 16).cuda().requires_grad_(True)
    b = torch.randn(Co, generator=g).cuda().requires_grad_(True)
    gy = torch.randn(B * T, Co, generator=g).cuda()
    outs = []
    for enabled in (True, False):
        ops.TC_ENABLED = enabled
        _lib.COUNTS.clear()
        for t in (x, w, b):
            t.grad = None
        z, st = ops.conv_gemm(x, w, b, B, T, want_stats=True)
        z.backward(gy)
        assert (_lib.COUNTS.get("tn_gemm_tc", 0) > 0) == enabled
        outs.append((z.detach().clone(), st.clone(), x.grad.clone(), w.grad.clone(), b.grad.clone()))
    ops.TC_ENABLED = True
    names = ("z", "stats", "dx", "dw", "db")
    for n, a, c in zip(names, *outs):
        # forward / dgrad run fp32-equivalent 3xTF32; the weight gradient uses plain TF32 operands
        assert rel(a, c) < (2e-3 if n == "dw" else 1e-5), n


@pytest.mark.parametrize("R,Ci,Co", [(19264, 256, 256), (1000, 256, 256), (2000, 256, 1536), (3000, 1536, 128), (777, 128, 1536),
                                      (5000, 512, 512), (64, 64, 128)])
def test_wgrad_tc_matches_fp64(R, Ci, Co):
    from titanet_b200._lib import call, ptr
    g = torch.Generator().manual_seed(R + Ci + Co)
    dz = torch.randn(R, Co, generator=g)
    u = torch.randn(R, Ci, generator=g)
    ref = dz.double().t() @ u.double()
    dw = torch.zeros(Co, Ci, device="cuda")
    dz_d, u_d = dz.cuda(), u.cuda()
    call("tn_wgrad_tc", ptr(dz_d), ptr(u_d), ptr(dw), R, Ci, Co)
    call("tn_wgrad_tc", ptr(dz_d), ptr(u_d), ptr(dw), R, Ci, Co)        # accumulates
    torch.cuda.synchronize()
    assert rel(dw, 2 * ref) < 2e-3, "plain-TF32 operands: ~1e-3"
    db = torch.zeros(Co, device="cuda")
    call("tn_colsum", ptr(dz_d), ptr(db), R, Co)
    assert rel(db, dz.double().sum(0)) < 1e-5


@pytest.mark.parametrize("B,T,C,Co,K,lazy,p", [(8, 301, 256, 256, 3, True, 0.0), (5, 77, 128, 256, 7, True, 0.0), (3, 120, 256, 128, 11, True, 0.0),
                                                (8, 301, 256, 256, 3, False, 0.0), (4, 100, 128, 128, 1, True, 0.0), (6, 301, 256, 256, 3, True, 0.1)])
def test_fused_dgrad_depthwise_backward(B, T, C, Co, K, lazy, p):
    """DwPw (depthwise-separable block): fused tensor-core backward == unfused kernels == fp64 torch."""
    import torch.nn.functional as F
    from titanet_b200 import _ops as ops
    g = torch.Generator().manual_seed(B * T + C + K)
    R = B * T
    z = torch.randn(R, C, generator=g)
    sc, sh = 0.5 + torch.rand(C, generator=g), 0.3 * torch.randn(C, generator=g)
    dw_w, dw_b = torch.randn(C, 1, K, generator=g) / math.sqrt(K), torch.randn(C, generator=g)
    pw_w, pw_b = torch.randn(Co, C, 1, generator=g) / math.sqrt(C), torch.randn(Co, generator=g)
    gy = torch.randn(R, Co, generator=g)
    seed = torch.tensor([77], dtype=torch.int64, device="cuda")
    res = []
    for fused in (True, False):
        ops.TC_FUSE_DWBWD = fused
        t = [x.clone().cuda().requires_grad_(True) for x in (z, sc, sh, dw_w, dw_b, pw_w, pw_b)]
        zo, st = ops.DwPw.apply(t[0], t[1] if lazy else None, t[2] if lazy else None, t[3], t[4], t[5], t[6], seed if p > 0 else None,
                                True, p, 5, B, T, True)
        (zo * gy.cuda()).sum().backward()
        res.append([zo.detach()] + [x.grad for x in t if x.grad is not None])
    ops.TC_FUSE_DWBWD = True
    for a, b in zip(*res):
        assert rel(a, b) < 2e-3           # weight gradients use plain-TF32 operands in both paths; the rest agrees to ~1e-5
    for i in (0, 1):                      # z_out and dz_prev: 3xTF32 in both paths
        assert rel(res[0][i], res[1][i]) < 2e-5
    if p == 0.0:
        zr, scr, shr, dwr, dbr, pwr, pbr = (x.double().clone().requires_grad_(True) for x in (z, sc, sh, dw_w, dw_b, pw_w, pw_b))
        zz = zr.view(B, T, C).permute(0, 2, 1)
        a = torch.relu(zz * scr.view(1, -1, 1) + shr.view(1, -1, 1)) if lazy else zz
        u = F.conv1d(F.pad(a, (K // 2, K // 2)), dwr, dbr, groups=C)
        zo_r = F.conv1d(u, pwr, pbr).permute(0, 2, 1).reshape(R, Co)
        (zo_r * gy.double()).sum().backward()
        ref = [zo_r.detach(), zr.grad] + ([scr.grad, shr.grad] if lazy else []) + [dwr.grad, dbr.grad, pwr.grad, pbr.grad]
        tols = [1e-5, 2e-5] + ([1e-4, 1e-4] if lazy else []) + [1e-4, 1e-4, 2e-3, 1e-4]
        for a_, r_, tol in zip(res[0], ref, tols):
            assert rel(a_, r_) < tol


def test_batched_split_writes_the_planes_the_gemms_read():
    """tn_split_tf32_batch (one tile list over all weights, transposed weights through shared memory, only the planes each
    orientation's GEMMs read) against the stand-alone tn_split_tf32 on ragged shapes: plane 0 (tf32 hi) always, plane 3
    (scaled fp16 correction rows) for the forward orientation, plane 2 (bf16 correction rows) for the transposed one."""
    from titanet_b200 import _ops as ops
    from titanet_b200._lib import call, ptr
    g = torch.Generator().manual_seed(11)
    ws = [torch.nn.Parameter((torch.randn(co, ci, 1, generator=g) / math.sqrt(ci)).cuda()) for co, ci in ((256, 256), (128, 1536), (1536, 128), (384, 96), (128, 32))]
    cache = ops.SplitCache(ws)
    cache.buf.fill_(float("nan"))
    cache.refresh()
    torch.cuda.synchronize()
    for w in ws:
        fwd, bwd = cache.lookup(w)
        Co, Ci = w.shape[0], w.shape[1]
        ref_f = torch.empty(ops.WS_PLANES, Co, Ci, device="cuda")
        ref_b = torch.empty(ops.WS_PLANES, Ci, Co, device="cuda")
        call("tn_split_tf32", ptr(w), ptr(ref_f), Co, Ci, 0)
        call("tn_split_tf32", ptr(w), ptr(ref_b), Ci, Co, 1)
        torch.cuda.synchronize()
        assert torch.equal(fwd[0], ref_f[0]) and torch.equal(fwd[3].view(torch.int32), ref_f[3].view(torch.int32))
        assert torch.equal(bwd[0], ref_b[0]) and torch.equal(bwd[2].view(torch.int32), ref_b[2].view(torch.int32))
        assert torch.equal(ref_b[0], ref_f[0].t())
