"""Per-kernel parity tests (GPU): every libtitanet_sm100 kernel group against a plain
torch fp64 CPU restatement of the same op, forward and backward, through the C ABI
(ctypes) exactly as the product path calls it."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import titanet_oracle as O  # noqa: E402  (checker only)


def dev():
    return torch.device("cuda:0")


def rel(a, b, floor=1e-30):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(floor))


def nwc(x):  # [B,C,T] -> [B*T, C]
    B, C, T = x.shape
    return x.permute(0, 2, 1).reshape(B * T, C).contiguous()


def ncw(x, B, T):  # [B*T, C] -> [B,C,T]
    return x.reshape(B, T, -1).permute(0, 2, 1).contiguous()


@pytest.fixture(scope="module")
def ops():
    from titanet_b200 import _ops
    return _ops


def test_library_loaded_and_device(ops):
    from titanet_b200._lib import LIB
    LIB.load()
    LIB.call("tn_device_check")


def test_transpose_roundtrip(ops):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 80, 101, generator=g)
    y = ops.ncw_to_nwc(x.to(dev()))
    assert torch.equal(y.cpu(), x.permute(0, 2, 1).contiguous())
    assert torch.equal(ops.nwc_to_ncw(y).cpu(), x)


@pytest.mark.parametrize("B,T,Ci,Co,K", [(2, 37, 80, 64, 3), (3, 50, 64, 64, 1), (2, 33, 48, 251, 1), (1, 70, 16, 96, 5),
                                         (4, 1, 3072, 192, 1), (300, 1, 3072, 192, 1), (64, 1, 192, 251, 1), (70, 1, 333, 100, 1),
                                         (1, 600, 64, 96, 1)])
def test_conv_gemm_fwd_bwd(ops, B, T, Ci, Co, K):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Ci, T, generator=g, dtype=torch.float64)
    w = torch.randn(Co, Ci, K, generator=g, dtype=torch.float64) / math.sqrt(Ci * K)
    b = torch.randn(Co, generator=g, dtype=torch.float64)
    gy = torch.randn(B, Co, T, generator=g, dtype=torch.float64)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    yr = O.conv1d_same(xr, wr, br)
    yr.backward(gy)
    xg = nwc(x).float().to(dev()).requires_grad_(True)
    wg = w.float().to(dev()).requires_grad_(True)
    bg = b.float().to(dev()).requires_grad_(True)
    z, stats = ops.conv_gemm(xg, wg, bg, B, T, want_stats=Co % 4 == 0)     # BatchNorm'd channel counts are multiples of 4
    assert rel(ncw(z, B, T), yr) < 1e-5
    if stats is not None:
        s1 = yr.sum(dim=(0, 2))
        s2 = (yr ** 2).sum(dim=(0, 2))
        assert rel(stats[:Co], s1, floor=1e-3) < 1e-5 and rel(stats[Co:], s2) < 1e-5
    z.backward(nwc(gy).float().to(dev()))
    assert rel(ncw(xg.grad, B, T), xr.grad) < 1e-5
    assert rel(wg.grad, wr.grad) < 1e-5
    assert rel(bg.grad, br.grad) < 1e-5


def test_conv_gemm_tanh_and_stats_grad(ops):
    """tanh epilogue and the gradient through the statistics output."""
    B, T, Ci, Co = 2, 41, 32, 48
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B * T, Ci, generator=g, dtype=torch.float64)
    w = torch.randn(Co, Ci, generator=g, dtype=torch.float64) / math.sqrt(Ci)
    b = torch.randn(Co, generator=g, dtype=torch.float64)
    cs = torch.randn(2 * Co, generator=g, dtype=torch.float64)
    gy = torch.randn(B * T, Co, generator=g, dtype=torch.float64)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    hr = torch.tanh(xr @ wr.t() + br)
    (hr * gy).sum().backward()
    xg, wg, bg = (t.float().to(dev()).requires_grad_(True) for t in (x, w, b))
    h, _ = ops.conv_gemm(xg, wg, bg, B, T, tanh=True)
    (h * gy.float().to(dev())).sum().backward()
    assert rel(h, hr) < 1e-5 and rel(xg.grad, xr.grad) < 2e-5 and rel(wg.grad, wr.grad) < 2e-5 and rel(bg.grad, br.grad) < 2e-5
    # statistics gradient
    xr2, wr2 = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    zr = xr2 @ wr2.t() + b
    lr = (torch.cat([zr.sum(0), (zr ** 2).sum(0)]) * cs).sum() + (zr * gy).sum()
    lr.backward()
    xg2, wg2 = x.float().to(dev()).requires_grad_(True), w.float().to(dev()).requires_grad_(True)
    z, st = ops.conv_gemm(xg2, wg2, b.float().to(dev()), B, T, want_stats=True)
    l = (st * cs.to(dev())).sum() + (z * gy.float().to(dev())).sum()
    l.backward()
    assert rel(xg2.grad, xr2.grad) < 2e-5 and rel(wg2.grad, wr2.grad) < 2e-5


def lazy_ref(z, scale, shift, relu):
    v = z * scale.view(1, -1, 1) + shift.view(1, -1, 1)
    return torch.relu(v) if relu else v


@pytest.mark.parametrize("K,C,T,B,lazy", [(3, 64, 50, 3, True), (7, 32, 33, 2, True), (11, 32, 64, 2, True), (3, 256, 101, 2, False),
                                          (7, 128, 45, 2, True), (11, 256, 70, 3, True), (9, 384, 37, 2, False), (15, 132, 40, 1, True),
                                          (1, 1536, 9, 2, True), (15, 16, 40, 1, True)])
def test_depthwise_fwd_bwd(ops, K, C, T, B, lazy):
    g = torch.Generator().manual_seed(3)
    z = torch.randn(B, C, T, generator=g, dtype=torch.float64)
    sc = (0.5 + torch.rand(C, generator=g, dtype=torch.float64))
    sh = 0.3 * torch.randn(C, generator=g, dtype=torch.float64)
    w = torch.randn(C, 1, K, generator=g, dtype=torch.float64) / math.sqrt(K)
    b = torch.randn(C, generator=g, dtype=torch.float64)
    gu = torch.randn(B, C, T, generator=g, dtype=torch.float64)
    zr, scr, shr, wr, br = (t.clone().requires_grad_(True) for t in (z, sc, sh, w, b))
    a = lazy_ref(zr, scr, shr, True) if lazy else zr
    ur = O.conv1d_same(a, wr, br, groups=C)
    ur.backward(gu)
    zg = nwc(z).float().to(dev()).requires_grad_(True)
    scg, shg, wg, bg = (t.float().to(dev()).requires_grad_(True) for t in (sc, sh, w, b))
    u = ops.Depthwise.apply(zg, scg if lazy else None, shg if lazy else None, wg, bg, None, True, 0.0, 0, B, T)
    assert rel(ncw(u, B, T), ur) < 1e-5
    u.backward(nwc(gu).float().to(dev()))
    assert rel(ncw(zg.grad, B, T), zr.grad) < 2e-5
    assert rel(wg.grad, wr.grad) < 2e-5 and rel(bg.grad, br.grad) < 2e-5
    if lazy:
        assert rel(scg.grad, scr.grad) < 2e-5 and rel(shg.grad, shr.grad) < 2e-5


@pytest.mark.parametrize("training", [True, False])
def test_bn_fold_act_chain(ops, training):
    """conv-GEMM -> BatchNorm (folded) -> ReLU, against F.batch_norm, incl. running stats."""
    B, T, Ci, Co = 3, 29, 32, 64
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, Ci, T, generator=g, dtype=torch.float64)
    w = torch.randn(Co, Ci, 1, generator=g, dtype=torch.float64) / math.sqrt(Ci)
    b = torch.randn(Co, generator=g, dtype=torch.float64)
    gamma = 0.5 + torch.rand(Co, generator=g, dtype=torch.float64)
    beta = 0.2 * torch.randn(Co, generator=g, dtype=torch.float64)
    rm = 0.1 * torch.randn(Co, generator=g, dtype=torch.float64)
    rv = 0.5 + torch.rand(Co, generator=g, dtype=torch.float64)
    gy = torch.randn(B, Co, T, generator=g, dtype=torch.float64)
    xr, wr, br, gr, ber = (t.clone().requires_grad_(True) for t in (x, w, b, gamma, beta))
    rm_r, rv_r = rm.clone(), rv.clone()
    yr = torch.relu(F.batch_norm(F.conv1d(xr, wr, br), rm_r, rv_r, gr, ber, training=training, momentum=0.1, eps=1e-5))
    yr.backward(gy)
    bn = torch.nn.BatchNorm1d(Co).to(dev())
    with torch.no_grad():
        bn.weight.copy_(gamma.float()); bn.bias.copy_(beta.float())
        bn.running_mean.copy_(rm.float()); bn.running_var.copy_(rv.float())
    bn.train(training)
    xg = nwc(x).float().to(dev()).requires_grad_(True)
    wg, bg = (t.float().to(dev()).requires_grad_(True) for t in (w, b))
    z, st = ops.conv_gemm(xg, wg, bg, B, T, want_stats=training)
    sc, sh = ops.bn_fold(st, bn, float(B * T))
    y = ops.Act.apply(z, sc, sh, None, True, 0.0, 0)
    assert rel(ncw(y, B, T), yr) < 1e-5
    y.backward(nwc(gy).float().to(dev()))
    assert rel(ncw(xg.grad, B, T), xr.grad) < 5e-5
    assert rel(wg.grad, wr.grad) < 5e-5
    assert rel(bn.weight.grad, gr.grad) < 5e-5 and rel(bn.bias.grad, ber.grad) < 5e-5
    if training:
        assert float(bg.grad.abs().max()) < 1e-4 * float(wr.grad.abs().max())   # mathematically zero: rounding noise only
        assert rel(bn.running_mean, rm_r) < 1e-5 and rel(bn.running_var, rv_r) < 1e-5
        assert int(bn.num_batches_tracked) == 1
    else:
        assert rel(bg.grad, br.grad) < 5e-5
        assert int(bn.num_batches_tracked) == 0


@pytest.mark.parametrize("R,C", [(64, 3072), (1203, 1536), (37, 80), (500, 192), (3000, 256), (9, 4)])
def test_act_two_consumers_and_colsum_slab_tiling(ops, R, C):
    """tn_act_fwd / tn_act_bwd2 (two gradients summed on load) / tn_colsum over the slab tiling: wide tensors (several 256-channel
    slabs), few rows, channel counts that do not fill a slab, row counts that are not a multiple of the unroll."""
    from titanet_b200._lib import call, ptr
    g = torch.Generator().manual_seed(R + C)
    z = torch.randn(R, C, generator=g, dtype=torch.float64)
    sc = 0.5 + torch.rand(C, generator=g, dtype=torch.float64)
    sh = 0.3 * torch.randn(C, generator=g, dtype=torch.float64)
    d1, d2 = torch.randn(R, C, generator=g, dtype=torch.float64), torch.randn(R, C, generator=g, dtype=torch.float64)
    zr, scr, shr = (t.clone().requires_grad_(True) for t in (z, sc, sh))
    yr = torch.relu(zr * scr + shr)
    (yr * d1).sum().backward(retain_graph=True)
    one = [t.grad.clone() for t in (zr, scr, shr)]
    (yr * d2).sum().backward()
    both = [t.grad.clone() for t in (zr, scr, shr)]
    zg, scg, shg = (t.float().to(dev()).requires_grad_(True) for t in (z, sc, sh))
    ya, yb = ops.Act2.apply(zg, scg, shg, None, True, 0.0, 0)
    assert ya.data_ptr() == yb.data_ptr() and rel(ya, yr) < 1e-6
    ((ya * d1.float().to(dev())).sum() + (yb * d2.float().to(dev())).sum()).backward()
    for got, want in zip((zg.grad, scg.grad, shg.grad), both):
        assert rel(got, want) < 2e-5
    zg.grad = scg.grad = shg.grad = None
    ya, yb = ops.Act2.apply(zg, scg, shg, None, True, 0.0, 0)
    (yb * d1.float().to(dev())).sum().backward()              # only the second alias is used
    for got, want in zip((zg.grad, scg.grad, shg.grad), one):
        assert rel(got, want) < 2e-5
    out = torch.zeros(C, device=dev())
    call("tn_colsum", ptr(zg.detach()), ptr(out), R, C)
    assert rel(out, z.float().double().sum(0)) < 2e-5


def test_colstats_bn_over_batch(ops):
    B, C = 6, 96
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, C, generator=g, dtype=torch.float64)
    gamma = 0.5 + torch.rand(C, generator=g, dtype=torch.float64)
    beta = torch.randn(C, generator=g, dtype=torch.float64)
    gy = torch.randn(B, C, generator=g, dtype=torch.float64)
    xr, gr, ber = (t.clone().requires_grad_(True) for t in (x, gamma, beta))
    yr = F.batch_norm(xr, None, None, gr, ber, training=True)
    yr.backward(gy)
    bn = torch.nn.BatchNorm1d(C).to(dev()).train()
    with torch.no_grad():
        bn.weight.copy_(gamma.float()); bn.bias.copy_(beta.float())
    xg = x.float().to(dev()).requires_grad_(True)
    st = ops.ColStats.apply(xg)
    sc, sh = ops.bn_fold(st, bn, float(B))
    y = ops.Act.apply(xg, sc, sh, None, False, 0.0, 0)
    assert rel(y, yr) < 1e-5
    y.backward(gy.float().to(dev()))
    assert rel(xg.grad, xr.grad) < 5e-5 and rel(bn.weight.grad, gr.grad) < 5e-5 and rel(bn.bias.grad, ber.grad) < 5e-5


def test_se_tail_fwd_bwd(ops):
    B, T, C, Cr = 3, 37, 64, 4
    g = torch.Generator().manual_seed(6)
    mk = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    z3, s = mk(B, C, T), mk(B, C, T)
    sc3, sh3 = 0.5 + torch.rand(C, generator=g, dtype=torch.float64), 0.3 * mk(C)
    scs, shs = 0.5 + torch.rand(C, generator=g, dtype=torch.float64), 0.3 * mk(C)
    W1, W2 = mk(Cr, C) / math.sqrt(C), mk(C, Cr) / math.sqrt(Cr)
    gy = mk(B, C, T)
    ref = [t.clone().requires_grad_(True) for t in (z3, sc3, sh3, s, scs, shs, W1, W2)]
    a3 = lazy_ref(ref[0], ref[1], ref[2], True)
    gate = torch.sigmoid(torch.relu(a3.mean(dim=2) @ ref[6].t()) @ ref[7].t())
    outr = torch.relu(lazy_ref(ref[3], ref[4], ref[5], False) + a3 * gate.unsqueeze(-1))
    outr.backward(gy)
    gp = [nwc(z3), sc3, sh3, nwc(s), scs, shs, W1, W2]
    gp = [t.float().to(dev()).requires_grad_(True) for t in gp]
    out = ops.SETail.apply(*gp, None, 0.0, 1, 0.0, 2, B, T)
    assert rel(ncw(out, B, T), outr) < 1e-5
    out.backward(nwc(gy).float().to(dev()))
    assert rel(ncw(gp[0].grad, B, T), ref[0].grad) < 5e-5
    assert rel(ncw(gp[3].grad, B, T), ref[3].grad) < 5e-5
    for i in (1, 2, 4, 5, 6, 7):
        assert rel(gp[i].grad, ref[i].grad) < 5e-5, i


@pytest.mark.parametrize("B,T,C,Cr,p3,po", [(5, 301, 256, 16, 0.1, 0.1), (3, 77, 512, 32, 0.0, 0.0), (2, 150, 128, 8, 0.2, 0.0), (2, 801, 64, 8, 0.1, 0.1)])
def test_se_tail_fused_forward_equals_two_launch_path(ops, B, T, C, Cr, p3, po):
    """tn_se_tail_fwd (squeeze + excitation + tail as one cluster kernel, activated tile kept in shared memory) against
    tn_se_squeeze_excite + tn_tail_fwd: same arithmetic in the same order, so means, gates and outputs are bit-identical."""
    from titanet_b200._lib import LIB
    if not LIB.query("tn_se_tail_fwd_supported", T, C, Cr):
        pytest.skip("shape outside the fused kernel's plan")
    g = torch.Generator().manual_seed(B * T + C)
    mk = lambda *s: torch.randn(*s, generator=g).to(dev())
    z3, s = mk(B * T, C), mk(B * T, C)
    sc3, sh3 = (0.5 + torch.rand(C, generator=g)).to(dev()), 0.3 * mk(C)
    scs, shs = (0.5 + torch.rand(C, generator=g)).to(dev()), 0.3 * mk(C)
    W1, W2 = mk(Cr, C) / math.sqrt(C), mk(C, Cr) / math.sqrt(Cr)
    seed = torch.tensor([12345], dtype=torch.int64, device=dev())
    outs = []
    for fused in (True, False):
        old = ops.FUSE_SE_TAIL
        ops.FUSE_SE_TAIL = fused
        try:
            out = ops.SETail.apply(z3, sc3, sh3, s, scs, shs, W1, W2, seed if (p3 > 0 or po > 0) else None, p3, 3, po, 4, B, T)
        finally:
            ops.FUSE_SE_TAIL = old
        outs.append(out.clone())
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    assert float(outs[0].abs().max()) > 0


def test_asp_pool_fwd_bwd(ops):
    B, T, D = 3, 45, 96
    g = torch.Generator().manual_seed(7)
    e = 2.0 * torch.randn(B, D, T, generator=g, dtype=torch.float64)
    x = torch.randn(B, D, T, generator=g, dtype=torch.float64)
    x[0, :5] = 0.25                      # constant channels: variance clamps at eps
    gy = torch.randn(B, 2 * D, generator=g, dtype=torch.float64)
    er, xr = e.clone().requires_grad_(True), x.clone().requires_grad_(True)
    a = torch.softmax(er, dim=2)
    mu = (a * xr).sum(2)
    pr = torch.cat([mu, torch.sqrt(((a * xr ** 2).sum(2) - mu ** 2).clamp(min=1e-6))], dim=1)
    pr.backward(gy)
    eg, xg = nwc(e).float().to(dev()).requires_grad_(True), nwc(x).float().to(dev()).requires_grad_(True)
    p = ops.ASPPool.apply(eg, xg, B, T, 1e-6)
    assert rel(p, pr) < 1e-5
    p.backward(gy.float().to(dev()))
    mask = torch.ones(B, D, T, dtype=torch.bool)
    mask[0, :5] = False                  # clamp boundary: fp32/fp64 may sit on different sides
    assert rel(ncw(eg.grad, B, T)[mask], er.grad[mask]) < 5e-5
    assert rel(ncw(xg.grad, B, T)[mask], xr.grad[mask]) < 5e-5


def test_l2norm_and_ce(ops):
    B, E, Cn = 5, 48, 251
    g = torch.Generator().manual_seed(8)
    x = torch.randn(B, E, generator=g, dtype=torch.float64)
    gy = torch.randn(B, E, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    F.normalize(xr, p=2, dim=1).backward(gy)
    xg = x.float().to(dev()).requires_grad_(True)
    y, n = ops.L2Norm.apply(xg, 1e-12)
    y.backward(gy.float().to(dev()))
    assert rel(y, F.normalize(x, dim=1)) < 1e-6 and rel(n, x.norm(dim=1)) < 1e-6 and rel(xg.grad, xr.grad) < 1e-5
    logits = 3 * torch.randn(B, Cn, generator=g, dtype=torch.float64)
    tgt = torch.randint(0, Cn, (B,), generator=g)
    lr = logits.clone().requires_grad_(True)
    loss_r = F.cross_entropy(lr, tgt)
    (2.5 * loss_r).backward()
    lg = logits.float().to(dev()).requires_grad_(True)
    loss, preds = ops.CrossEntropy.apply(lg, tgt.to(dev()))
    (2.5 * loss).backward()
    assert rel(loss, loss_r) < 1e-6 and torch.equal(preds.cpu(), logits.argmax(1)) and rel(lg.grad, lr.grad) < 1e-5


@pytest.mark.parametrize("scale,m1,m2,m3", [(30.0, 1.0, 0.2, 0.0), (64.0, 1.0, 0.0, 0.2), (None, 1.0, 0.2, 0.0), (None, 3.0, 0.0, 0.0)])
def test_margin_loss(ops, scale, m1, m2, m3):
    B, E, Cn = 6, 48, 10
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, E, generator=g, dtype=torch.float64)
    w = torch.randn(Cn, E, generator=g, dtype=torch.float64)
    tgt = torch.randint(0, Cn, (B,), generator=g)
    sd = {"loss_function.fc.weight": w.clone().requires_grad_(True)}
    xr = x.clone().requires_grad_(True)
    xh_r, preds_r, loss_r = O.angular_margin_head(sd, xr, tgt, scale, m1, m2, m3)
    loss_r.backward()
    wn = sd["loss_function.fc.weight"]
    from titanet_b200 import losses
    head = losses.AngularMarginLoss(E, Cn, scale=scale, m1=m1, m2=m2, m3=m3).to(dev())
    with torch.no_grad():
        head.fc.weight.copy_(w.float())
    xg = x.float().to(dev()).requires_grad_(True)
    xh, preds, loss = head(xg, tgt.to(dev()))
    loss.backward()
    assert rel(head.fc.weight, wn) < 1e-6                      # in-place renormalisation side effect
    assert rel(xh, xh_r) < 1e-6 and torch.equal(preds.cpu(), preds_r)
    assert rel(loss, loss_r) < 1e-5
    assert rel(xg.grad, xr.grad) < 1e-4 and rel(head.fc.weight.grad, wn.grad) < 1e-4


def test_mel_against_golden_and_oracle(golden_dir):
    from titanet_b200 import transforms
    from cases import mel_inputs
    gold = np.load(os.path.join(golden_dir, "mel_1s.npz"))
    mel = transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=160, n_mels=80, specaugment_probability=0.0)
    wave, _ = O.synthetic_batch(2, seconds=1.0, seed=42)
    out = mel.batch(wave.to(dev()))
    assert out.shape == (2, 80, 101)
    assert rel(out, gold["mel"]) < 2e-5
    out_cl = mel.batch(wave.to(dev()), channels_last=True)
    assert torch.equal(out_cl.permute(0, 2, 1).contiguous(), out)
    for k, w in mel_inputs().items():
        ex = mel({"waveform": w.view(1, -1), "sample_rate": 16000})
        assert ex["spectrogram"].device.type == "cpu" and ex["spectrogram"].shape == (1,) + gold[k].shape
        assert rel(ex["spectrogram"][0], gold[k]) < 2e-5, k
    # ragged batch == per-utterance transform + collate zero padding (datasets.py:48-73)
    g = torch.Generator().manual_seed(11)
    lens = [16000, 12345, 4000, 48000]
    waves = [0.1 * torch.randn(n, generator=g) for n in lens]
    batch = torch.zeros(len(lens), max(lens))
    for i, w in enumerate(waves):
        batch[i, : len(w)] = w
    got = mel.batch(batch.to(dev()), torch.tensor(lens).to(dev()))
    want, _ = O.collate_pad([O.mel_spectrogram(w.view(1, -1)) for w in waves])
    assert got.shape == want.shape and rel(got, want) < 2e-5
    assert float(got[2, :, 26:].abs().max()) == 0.0


def test_mel_specaugment_against_golden_and_oracle(golden_dir):
    """SpecAugment inside the mel kernel (tn_mel_specaug_fwd): the example-dict call under the reference's seeds vs the
    reference-generated golden; a ragged batch with per-utterance draws vs the oracle + collate zero padding."""
    import random
    from titanet_b200 import transforms
    from cases import SPECAUG_CASES, SPECAUG_KW, specaug_wave
    gold = np.load(os.path.join(golden_dir, "mel_specaug.npz"))
    for i, (samples, seed) in enumerate(SPECAUG_CASES):
        kw = {}
        if i == len(SPECAUG_CASES) - 1:
            kw = dict(specaugment_freq_mask_num=SPECAUG_KW["freq_mask_num"], specaugment_time_mask_num=SPECAUG_KW["time_mask_num"])
        mel = transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=160, n_mels=80, **kw)   # probability 1.0
        random.seed(seed)
        torch.manual_seed(seed)
        ex = mel({"waveform": specaug_wave(samples, seed), "sample_rate": 16000})
        want = torch.from_numpy(gold[f"specaug_{samples}_{seed}"])
        assert ex["spectrogram"].shape == (1,) + want.shape
        assert rel(ex["spectrogram"][0], want) < 5e-5, (samples, seed)     # fp32 radix-2 FFT + sqrt/interpolate vs torch (measured 2.1e-5)
        assert torch.equal(ex["spectrogram"][0] == 0, want == 0)                  # identical masks
    # ragged batch: utterance 1 not augmented, utterance 3 stretched only, others stretched + masked
    mel = transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=160, n_mels=80)
    g = torch.Generator().manual_seed(13)
    lens = [16000, 12345, 48000, 30001]
    waves = [0.1 * torch.randn(n, generator=g) for n in lens]
    batch = torch.zeros(len(lens), max(lens))
    for i, w in enumerate(waves):
        batch[i, : len(w)] = w
    SD = transforms.SpecAugmentDraw
    fr = [mel.n_frames(n) for n in lens]
    draws = [SD(0.95, transforms.stretched_frames(fr[0], 0.95), [(3, 20)], [(10, 25)]), None,
             SD(1.05, transforms.stretched_frames(fr[2], 1.05), [(70, 80), (0, 2)], [(0, 7)]),
             SD(1.0123456789, transforms.stretched_frames(fr[3], 1.0123456789))]
    got = mel.batch(batch.to(dev()), torch.tensor(lens).to(dev()), augment=draws)
    each = [O.mel_spectrogram(w.view(1, -1)) if d is None else
            O.mel_spectrogram_specaugment(w.view(1, -1), d.rate, d.freq_masks, d.time_masks) for w, d in zip(waves, draws)]
    want, _ = O.collate_pad(each)
    assert got.shape == want.shape and rel(got, want) < 5e-5
    got_cl = mel.batch(batch.to(dev()), torch.tensor(lens).to(dev()), channels_last=True, augment=draws)
    assert torch.equal(got_cl.permute(0, 2, 1).contiguous(), got)
    auto = mel.batch(batch.to(dev()), torch.tensor(lens).to(dev()), augment=True)       # host draws, probability 1.0
    assert auto.shape[:2] == (4, 80) and bool(torch.isfinite(auto).all())
    assert float((auto == 0).float().mean()) > 0.02                                      # some masked / padded cells


def test_dropout_statistics_and_mask_replay(ops):
    """dropout>0 cannot match torch's RNG stream; check keep-rate, 1/(1-p) scaling, and that
    backward regenerates the forward mask."""
    R, C, p = 4096, 64, 0.1
    z = torch.ones(R, C, device=dev(), requires_grad=True)
    sc, sh = torch.ones(C, device=dev()), torch.zeros(C, device=dev())
    state = torch.tensor([1234], dtype=torch.int64, device=dev())
    seed = ops.seed_next(state)
    y = ops.Act.apply(z, sc, sh, seed, True, p, 3)
    keep = (y > 0).float().mean().item()
    assert abs(keep - (1 - p)) < 0.01
    vals = torch.unique(y)
    assert vals.numel() == 2 and abs(float(vals.max()) - 1 / (1 - p)) < 1e-6
    y.sum().backward()
    assert torch.equal((z.grad > 0), (y > 0)) and abs(float(z.grad.max()) - 1 / (1 - p)) < 1e-6
    y2 = ops.Act.apply(z.detach(), sc, sh, seed, True, p, 4)       # another layer id -> another mask
    assert not torch.equal(y2 > 0, y > 0)
    seed2 = ops.seed_next(state)
    y3 = ops.Act.apply(z.detach(), sc, sh, seed2, True, p, 3)      # next step -> another mask
    assert not torch.equal(y3 > 0, y > 0)


def test_standalone_squeeze_excitation_matches_fp64():
    """modules.SqueezeExcitation.forward outside a MegaBlock (src/modules.py:173-189): [B, C, W] -> [B, C, W]."""
    from titanet_b200 import modules
    B, C, T = 5, 64, 77
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, C, T, generator=g)
    gy = torch.randn(B, C, T, generator=g)
    se = modules.SqueezeExcitation(C, reduction=16)
    W1, W2 = se.excitation[0].weight.detach().clone(), se.excitation[2].weight.detach().clone()
    xr, w1r, w2r = x.double().requires_grad_(True), W1.double().requires_grad_(True), W2.double().requires_grad_(True)
    gate = torch.sigmoid(torch.relu(xr.mean(dim=2) @ w1r.t()) @ w2r.t())
    yr = xr * gate.unsqueeze(-1)
    yr.backward(gy.double())
    se = se.to(dev())
    xg = x.to(dev()).requires_grad_(True)
    y = se(xg)
    y.backward(gy.to(dev()))
    assert y.shape == (B, C, T)
    assert rel(y, yr) < 1e-5
    assert rel(xg.grad, xr.grad) < 1e-4
    assert rel(se.excitation[0].weight.grad, w1r.grad) < 1e-4 and rel(se.excitation[2].weight.grad, w2r.grad) < 1e-4


def test_fused_adam_matches_torch_adam():
    """optim.FusedAdam == torch.optim.Adam (src/train.py:130-136) over several steps, incl. weight decay and an LR change."""
    from titanet_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(3)
    shapes = [(64, 32, 3), (64,), (7, 5), (1,), (1000, 33)]
    ref_p = [torch.randn(s, generator=g, dtype=torch.float64).requires_grad_(True) for s in shapes]
    our_p = [p.detach().float().to(dev()).requires_grad_(True) for p in ref_p]
    for wd in (0.0, 0.01):
        ref = torch.optim.Adam(ref_p, lr=1e-3, weight_decay=wd)
        ours = FusedAdam(our_p, lr=1e-3, weight_decay=wd)
        for it in range(4):
            if it == 2:
                for o in (ref, ours):
                    o.param_groups[0]["lr"] = 5e-4
            for pr, po in zip(ref_p, our_p):
                gr = torch.randn(pr.shape, generator=g, dtype=torch.float64)
                pr.grad = gr.clone()
                po.grad = gr.float().to(dev())
            ref.step()
            ours.step()
        for pr, po in zip(ref_p, our_p):
            assert rel(po, pr) < 2e-6
        assert rel(ours.state[our_p[0]]["exp_avg"], ref.state[ref_p[0]]["exp_avg"]) < 1e-5
        assert rel(ours.state[our_p[0]]["exp_avg_sq"], ref.state[ref_p[0]]["exp_avg_sq"]) < 1e-5
