"""Generate tests/golden/*.npz by running the UNMODIFIED reference here.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python oracle/make_golden.py``.  It imports the reference's own
``modules / models / losses / transforms`` from ``/root/reference/src``, loads the
deterministic synthetic weights of ``oracle.titanet_oracle.synth_state_dict`` with
``load_state_dict(strict=True)`` (which also pins the state_dict key schema), runs
the reference forward (+ backward) on seeded synthetic inputs and stores the
outputs.  Inputs and weights are NOT stored: tests regenerate them from the same
seeds.  TEST INFRASTRUCTURE.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get("TITANET_REFERENCE", "/root/reference/src")
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import titanet_oracle as O  # noqa: E402
import modules, models, losses, transforms  # noqa: E402,F401  (the reference)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(8)


def ref_mel(wave_1d: torch.Tensor) -> torch.Tensor:
    """The reference front end exactly as ``train.py:25-46`` configures it
    (get_transforms defaults: n_fft 512, win 25 ms, hop 10 ms, 80 mels; SpecAugment
    probability 0)."""
    t = transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=160, n_mels=80,
                                  specaugment_probability=0.0)
    return t({"waveform": wave_1d.view(1, -1), "sample_rate": 16000})["spectrogram"]


def build_ref(spec: O.TitaNetSpec, loss, n_classes, scale=None, margin=None, seed=42):
    lf = None
    if loss == "ce":
        lf = losses.CELoss(spec.emb, n_classes)
    elif loss == "arc":
        lf = losses.ArcFaceLoss(spec.emb, n_classes, scale=scale, margin=margin)
    elif loss == "cos":
        lf = losses.CosFaceLoss(spec.emb, n_classes, scale=scale, margin=margin)
    elif loss == "sphere":
        lf = losses.SphereFaceLoss(spec.emb, n_classes, scale=scale, margin=margin)
    m = models.TitaNet(spec.n_mels, spec.n_mega_blocks, spec.n_sub_blocks, spec.hidden, spec.enc_out,
                       spec.emb, spec.kernel, prolog_kernel_size=spec.prolog_kernel,
                       epilog_kernel_size=spec.epilog_kernel, attention_hidden_size=spec.attn_hidden,
                       se_reduction=spec.se_reduction, simple_pool=spec.simple_pool, loss_function=lf,
                       dropout=spec.dropout)
    sd = O.synth_state_dict(spec, loss, n_classes, seed=seed)
    assert list(m.state_dict().keys()) == list(sd.keys()), "state_dict key schema mismatch"
    m.load_state_dict(sd, strict=True)
    return m


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name + ".npz"),
                        **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                           for k, v in arrs.items()})
    print("wrote", name, len(arrs), "arrays")


def case_mel():
    wave, _ = O.synthetic_batch(2, seconds=1.0, seed=42)
    mels = torch.cat([ref_mel(w) for w in wave], dim=0)                 # [2, 80, 101]
    extra = {k: ref_mel(w)[0] for k, w in mel_inputs().items()}
    save("mel_1s", mel=mels, **extra)


def case_specaugment():
    """``MelSpectrogram.__call__`` with SpecAugment (src/transforms.py:165-201) under fixed ``random`` / ``torch`` seeds:
    the draws come out of the reference's own code, so the stored spectrograms pin the arithmetic AND the order /
    source of the random numbers."""
    import random
    from cases import SPECAUG_CASES, SPECAUG_KW, specaug_wave
    out = {}
    for i, (samples, seed) in enumerate(SPECAUG_CASES):
        kw = dict(specaugment_probability=1.0)
        if i == len(SPECAUG_CASES) - 1:
            kw.update(specaugment_freq_mask_num=SPECAUG_KW["freq_mask_num"], specaugment_time_mask_num=SPECAUG_KW["time_mask_num"])
        t = transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=160, n_mels=80, **kw)
        random.seed(seed)
        torch.manual_seed(seed)
        out[f"specaug_{samples}_{seed}"] = t({"waveform": specaug_wave(samples, seed), "sample_rate": 16000})["spectrogram"][0]
    save("mel_specaug", **out)


def case_cfg1():
    """BASELINE.json configs[0]: TitaNet-S eval forward, batch 2, 1 s synthetic waveform."""
    spec = O.TitaNetSpec.named("s", 17, dropout=0.1)
    m = build_ref(spec, None, 0).eval()
    wave, _ = O.synthetic_batch(2, seconds=1.0, seed=42)
    x = torch.cat([ref_mel(w) for w in wave], dim=0)
    with torch.no_grad():
        emb = m(x)
    save("cfg1_s17_eval", emb=emb, n_params=np.int64(m.get_n_params()))


from cases import TINY, TRAIN_CASES, ORACLE_ONLY_CASES, BIG_CASES, big_inputs, train_inputs, mel_inputs, eval_dx_inputs  # noqa: E402


def train_case(name, spec, loss, n_classes, B, T, scale=None, margin=None, full_grads=True, seed=42,
               input_grad=True):
    m = build_ref(spec, loss, n_classes, scale, margin, seed).train()
    x, y = train_inputs(spec, n_classes, B, T, seed)
    x.requires_grad_(input_grad)
    emb, preds, lval = m(x, speakers=y)
    lval.backward()
    out = dict(emb=emb, preds=preds, loss=lval)
    names = [k for k, _ in m.named_parameters()]
    gn = torch.stack([p.grad.norm() if p.grad is not None else torch.zeros(()) for _, p in m.named_parameters()])
    out["grad_norms"] = gn
    if full_grads:
        for k, p in m.named_parameters():
            out["grad:" + k] = p.grad
    else:
        for k, p in m.named_parameters():
            if p.numel() <= 2048:
                out["grad:" + k] = p.grad
    if input_grad:
        out["dx"] = x.grad
    sd_after = m.state_dict()
    for k in ("encoder.prolog.conv_block.1.running_mean", "encoder.prolog.conv_block.1.running_var",
              "decoder.linear.1.running_mean", "decoder.linear.1.running_var",
              "decoder.linear.1.num_batches_tracked"):
        out["buf:" + k] = sd_after[k]
    if loss != "ce":
        out["buf:loss_function.fc.weight"] = sd_after["loss_function.fc.weight"]
    save(name, **out)
    return names


def big_case(name, case):
    """TitaNet-M/10 and TitaNet-L/5 (ragged) train steps with the ArcFace head: embeddings, loss, predictions, input gradient,
    the gradients of every small parameter tensor and the norm of every gradient."""
    m = build_ref(case["spec"], case["loss"], case["nc"], case["scale"], case["margin"]).train()
    x, y, frames = big_inputs(case)
    x.requires_grad_(True)
    emb, preds, lval = m(x, speakers=y)
    lval.backward()
    out = dict(emb=emb, preds=preds, loss=lval, dx=x.grad, frames=frames)
    out["grad_norms"] = torch.stack([p.grad.norm() for _, p in m.named_parameters()])
    for k, p in m.named_parameters():
        if p.numel() <= 2048:
            out["grad:" + k] = p.grad
    out["buf:loss_function.fc.weight"] = m.state_dict()["loss_function.fc.weight"]
    save(name, **out)


def case_checkpoint():
    """A checkpoint written by the reference's own ``save_checkpoint`` (src/learn.py:180-201, extracted with ``ast`` because
    ``learn`` imports wandb / rich / matplotlib, which are not installed) after two ``torch.optim.Adam`` steps of the
    reference model (train.py:130-136: lr 1e-3, weight_decay 0; CosineAnnealingLR as train.py:138-144), plus the
    parameters the reference reaches after a THIRD step on the same batch: what a resumed run must reproduce."""
    import ast
    import tempfile
    src = open(os.path.join(REF, "learn.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "save_checkpoint")
    ns = {"os": os, "torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "learn.py", "exec"), ns)
    spec, loss, nc, B, T, scale, margin, _ = TRAIN_CASES["tiny_k3_arc"]
    m = build_ref(spec, loss, nc, scale, margin).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=0)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10)
    x, y = train_inputs(spec, nc, B, T)

    def step():
        _, _, lval = m(x, speakers=y)
        opt.zero_grad()
        lval.backward()
        opt.step()
        return float(lval)

    losses_seen = [step(), step()]
    sched.step()
    with tempfile.TemporaryDirectory() as d:
        ns["save_checkpoint"](2, d, m, opt, lr_scheduler=sched)
        blob = open(os.path.join(d, "epoch_2.pth"), "rb").read()
    with open(os.path.join(OUT, "ref_checkpoint_tiny_k3_arc.pth"), "wb") as f:
        f.write(blob)
    losses_seen.append(step())
    after = {"param:" + k: p.detach().clone() for k, p in m.named_parameters()}
    save("ref_checkpoint_tiny_k3_arc_step3", losses=np.asarray(losses_seen), lr=np.float64(opt.param_groups[0]["lr"]), **after)
    print("wrote ref_checkpoint_tiny_k3_arc.pth", len(blob), "bytes")


def case_fp64():
    """The gradient yard-stick of the GPU tests is the oracle evaluated in fp64; pin THAT against the reference modules run
    in fp64 (``model.double()``), train mode, CE and ArcFace heads."""
    for name, loss, scale, margin in (("tiny_k3_ce_fp64", "ce", None, None), ("tiny_k3_arc_fp64", "arc", 30, 0.2)):
        spec, nc, B, T = TINY["tiny_k3"], 10, 4, 50
        m = build_ref(spec, loss, nc, scale, margin).double().train()
        x, y = train_inputs(spec, nc, B, T)
        x = x.double().requires_grad_(True)
        emb, preds, lval = m(x, speakers=y)
        lval.backward()
        out = dict(emb=emb, preds=preds, loss=lval, dx=x.grad)
        for k, p in m.named_parameters():
            out["grad:" + k] = p.grad
        save(name, **out)


def case_eval_input_grad():
    """``utils.chart_dependencies`` (utils.py:451-468): eval-mode forward, backprop one
    sample's outputs, only that sample's input gradient is non-zero."""
    spec = TINY["tiny_k3"]
    m = build_ref(spec, None, 0).eval()
    x = eval_dx_inputs().requires_grad_(True)
    out = m(x)
    out[1].sum().backward()
    save("tiny_k3_eval_dx", emb=out, dx=x.grad)


if __name__ == "__main__":
    only = set(sys.argv[1:])            # optional: regenerate just the named train cases
    if only == {"specaug"}:
        case_specaugment()
        sys.exit(0)
    if only == {"fp64"}:
        case_fp64()
        sys.exit(0)
    if only == {"big"}:
        for name, case in BIG_CASES.items():
            big_case(name, case)
        sys.exit(0)
    if only == {"checkpoint"}:
        case_checkpoint()
        sys.exit(0)
    if not only:
        case_mel()
        case_specaugment()
        case_cfg1()
    for name, (spec, loss, nc, B, T, scale, margin, full) in {**TRAIN_CASES, **ORACLE_ONLY_CASES}.items():
        if only and name not in only:
            continue
        train_case(name, spec, loss, nc, B, T, scale if loss != "ce" else None,
                   margin if loss != "ce" else None, full)
    if not only:
        case_eval_input_grad()
        case_fp64()
        for name, case in BIG_CASES.items():
            big_case(name, case)
        case_checkpoint()
