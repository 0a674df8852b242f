"""Whole-path parity (GPU): titanet_b200 modules against (a) the committed golden vectors
produced by the reference itself and (b) the CPU oracle on the same seeded inputs.

Tolerances: the north star asks for embeddings and loss within 1e-3 relative of the
reference's fp32 CPU forward.  Gradients are judged against the fp64 oracle with the
fp32 oracle's own error as the yard-stick (train-mode BatchNorm amplifies fp32 rounding;
SURVEY.md §7 hard part 1)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import titanet_oracle as O  # noqa: E402  (checker only)
from cases import TRAIN_CASES, TINY, train_inputs, eval_dx_inputs  # noqa: E402


def rel(a, b, floor=1e-30):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(floor))


def build_model(spec, loss=None, n_classes=0, scale=None, margin=None, seed=42, device="cuda:0"):
    from titanet_b200 import losses, models
    lf = None
    if loss == "ce":
        lf = losses.CELoss(spec.emb, n_classes)
    elif loss == "arc":
        lf = losses.ArcFaceLoss(spec.emb, n_classes, scale=scale, margin=margin)
    elif loss == "cos":
        lf = losses.CosFaceLoss(spec.emb, n_classes, scale=scale, margin=margin)
    elif loss == "sphere":
        lf = losses.SphereFaceLoss(spec.emb, n_classes, scale=scale, margin=margin)
    m = models.TitaNet(spec.n_mels, spec.n_mega_blocks, spec.n_sub_blocks, spec.hidden, spec.enc_out, spec.emb, spec.kernel,
                       prolog_kernel_size=spec.prolog_kernel, epilog_kernel_size=spec.epilog_kernel,
                       attention_hidden_size=spec.attn_hidden, se_reduction=spec.se_reduction, simple_pool=spec.simple_pool,
                       loss_function=lf, dropout=spec.dropout)
    m.load_state_dict(O.synth_state_dict(spec, loss, n_classes, seed=seed), strict=True)
    return m.to(device)


def test_cfg1_eval_forward_matches_reference(golden_dir):
    """BASELINE.json configs[0]: TitaNet-S eval forward, batch 2, 1 s synthetic waveform,
    mel computed by the CUDA front end."""
    from titanet_b200 import transforms
    g = np.load(os.path.join(golden_dir, "cfg1_s17_eval.npz"))
    spec = O.TitaNetSpec.named("s", 17, dropout=0.1)
    model = build_model(spec).eval()
    assert int(model.get_n_params()) == int(g["n_params"])
    wave, _ = O.synthetic_batch(2, seconds=1.0, seed=42)
    mel = transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=160, n_mels=80, specaugment_probability=0.0)
    with torch.no_grad():
        emb = model(mel.batch(wave.cuda()))
    assert emb.shape == (2, 192)
    assert rel(emb, g["emb"]) < 1e-3
    assert torch.allclose(emb.norm(dim=1).cpu(), torch.ones(2), atol=1e-5)


@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_train_step_matches_reference(golden_dir, name):
    spec, loss, nc, B, T, scale, margin, full = TRAIN_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    model = build_model(spec, loss, nc, scale, margin).train()
    x, y = train_inputs(spec, nc, B, T)
    xg = x.cuda().requires_grad_(True)
    emb, preds, lval = model(xg, speakers=y.cuda())
    lval.backward()
    sd64 = O.synth_state_dict(spec, loss, nc, dtype=torch.float64)
    r64 = O.titanet_step(sd64, spec, x.double(), y, loss, scale=scale, margin=margin, input_grad=True)
    if spec.n_mega_blocks > 4:
        # S/17 at batch 4 (my stress case, not a BASELINE config): 72 train-mode BatchNorms over a tiny batch amplify
        # fp32 rounding so much that the reference's own fp32 forward sits ~6e-4 from its fp64 run.  Two independent
        # fp32 evaluations can therefore differ by ~1e-3; judge against fp64 with the reference's error as yard-stick.
        ref_err = rel(g["emb"], r64[0])
        assert rel(emb, r64[0]) <= max(1e-3, 2.0 * ref_err), "embeddings vs fp64 oracle"
        assert rel(emb, g["emb"]) < 2e-3, "embeddings vs reference"
    else:
        assert rel(emb, g["emb"]) < 1e-3, "embeddings vs reference"
    assert abs(float(lval) - float(g["loss"])) <= 1e-3 * abs(float(g["loss"])), "loss vs reference"
    assert np.array_equal(preds.cpu().numpy(), g["preds"])
    # gradients: fp64 oracle as truth, fp32 reference (golden) error as the yard-stick
    grads64, dx64 = r64[3], r64[5]
    gmax = max(float(v.abs().max()) for v in grads64.values())
    worst_ours, worst_ref = 0.0, 0.0
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        worst_ours = max(worst_ours, rel(p.grad, grads64[k], floor=1e-3 * gmax))
        if "grad:" + k in g.files:
            worst_ref = max(worst_ref, rel(g["grad:" + k], grads64[k], floor=1e-3 * gmax))
    assert worst_ours <= max(3.0 * worst_ref, 2e-3), (worst_ours, worst_ref)
    dx_ours, dx_ref = rel(xg.grad, dx64), rel(g["dx"], dx64)
    assert dx_ours <= max(3.0 * dx_ref, 2e-3), (dx_ours, dx_ref)
    # BatchNorm buffers and the ArcFace in-place weight renormalisation
    sd_after = model.state_dict()
    deep = spec.n_mega_blocks > 4
    for k in g.files:
        if k.startswith("buf:"):
            assert rel(sd_after[k[4:]], g[k]) < (2e-3 if deep else 1e-4), k


def test_eval_input_gradient_is_per_sample(golden_dir):
    """utils.chart_dependencies (utils.py:451-468): eval-mode, backprop one sample."""
    g = np.load(os.path.join(golden_dir, "tiny_k3_eval_dx.npz"))
    model = build_model(TINY["tiny_k3"]).eval()
    x = eval_dx_inputs().cuda().requires_grad_(True)
    emb = model(x)
    emb[1].sum().backward()
    assert rel(emb, g["emb"]) < 1e-4
    assert rel(x.grad, g["dx"]) < 1e-3
    assert bool((x.grad[0] == 0).all()) and bool((x.grad[2] == 0).all()) and bool((x.grad[1] != 0).any())


def test_cfg2_shape_against_oracle():
    """BASELINE.json configs[1] shape (S/17 + CE, 3 s utterances) at a batch the CPU oracle
    finishes in seconds; dropout 0 for parity."""
    from titanet_b200 import transforms
    spec = O.TitaNetSpec.named("s", 17)
    B = 16
    wave, labels = O.synthetic_batch(B, seconds=3.0, n_classes=251, seed=42)
    mel = transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=160, n_mels=80, specaugment_probability=0.0)
    model = build_model(spec, "ce", 251).train()
    x = mel.batch(wave.cuda())
    emb, preds, loss = model(x, speakers=labels.cuda())
    loss.backward()
    x_ref = torch.cat([O.mel_spectrogram(w.view(1, -1)) for w in wave])
    assert rel(x, x_ref) < 2e-5
    sd = O.synth_state_dict(spec, "ce", 251)
    emb_r, preds_r, loss_r, grads_r, _ = O.titanet_step(sd, spec, x_ref, labels, "ce")
    assert rel(emb, emb_r) < 1e-3
    assert abs(float(loss) - float(loss_r)) <= 1e-3 * abs(float(loss_r))
    assert torch.isfinite(loss) and all(torch.isfinite(p.grad).all() for p in model.parameters())


def test_dropout_training_runs_and_is_finite():
    spec = O.TitaNetSpec.named("s", 2, dropout=0.1)
    model = build_model(spec, "ce", 251).train()
    x, y = train_inputs(spec, 251, 4, 101)
    out1 = model(x.cuda(), speakers=y.cuda())
    out1[2].backward()
    out2 = model(x.cuda(), speakers=y.cuda())
    assert torch.isfinite(out1[2]) and torch.isfinite(out2[2])
    assert float(out1[2]) != float(out2[2])          # a fresh mask per forward
    model.eval()
    with torch.no_grad():
        e1, e2 = model(x.cuda()), model(x.cuda())
    # no dropout in eval mode (bitwise equality is not promised: reductions use float atomics)
    assert torch.allclose(e1, e2, rtol=0, atol=1e-6)
