"""Checkpoint compatibility + resume on the device.

``tests/golden/ref_checkpoint_tiny_k3_arc.pth`` was written by the REFERENCE's own ``save_checkpoint`` (src/learn.py:180-201,
run by oracle/make_golden.py:case_checkpoint) after two ``torch.optim.Adam`` steps of the reference model; the companion
``..._step3.npz`` holds the parameters the reference reaches after a third step on the same batch.  The CUDA model must load
the file (``strict=True``), ``FusedAdam`` must resume from the reference's optimizer state (moments AND step count), and the
resumed step must land where the reference's own third step lands."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import titanet_oracle as O  # noqa: E402,F401  (checker only)
from cases import TRAIN_CASES, train_inputs  # noqa: E402


def _load(golden_dir):
    ckpt = torch.load(os.path.join(golden_dir, "ref_checkpoint_tiny_k3_arc.pth"), map_location="cpu", weights_only=True)
    after = np.load(os.path.join(golden_dir, "ref_checkpoint_tiny_k3_arc_step3.npz"))
    return ckpt, after


def _model(ckpt):
    from titanet_b200 import losses, models
    spec, loss, nc, B, T, scale, margin, _ = TRAIN_CASES["tiny_k3_arc"]
    lf = losses.ArcFaceLoss(spec.emb, nc, scale=scale, margin=margin)
    m = models.TitaNet(spec.n_mels, spec.n_mega_blocks, spec.n_sub_blocks, spec.hidden, spec.enc_out, spec.emb, spec.kernel,
                       attention_hidden_size=spec.attn_hidden, se_reduction=spec.se_reduction, loss_function=lf, dropout=spec.dropout)
    res = m.load_state_dict(ckpt["model"], strict=True)          # the reference's key schema, loss head included
    assert not res.missing_keys and not res.unexpected_keys
    return m.cuda().train(), train_inputs(spec, nc, B, T)


def test_resume_from_reference_checkpoint(golden_dir):
    from titanet_b200.optim import FusedAdam
    ckpt, after = _load(golden_dir)
    assert set(ckpt) == {"model", "optimizer", "lr_scheduler", "epoch"} and ckpt["epoch"] == 2
    model, (x, y) = _model(ckpt)
    opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=0)
    opt.load_state_dict(ckpt["optimizer"])
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10)
    sched.load_state_dict(ckpt["lr_scheduler"])
    assert abs(opt.param_groups[0]["lr"] - float(after["lr"])) < 1e-12
    _, _, loss = model(x.cuda(), speakers=y.cuda())
    # after the forward: the ArcFace head has renormalised fc.weight in place (src/losses.py:86), as in the reference's step
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    assert abs(float(loss) - float(after["losses"][2])) <= 1e-3 * abs(float(after["losses"][2]))     # the loss of the reference's 3rd step
    opt.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    opt.step()
    torch.cuda.synchronize()
    # (1) the resumed FusedAdam step == torch.optim.Adam resumed from the same file, on the same gradients
    clones = [torch.nn.Parameter(before[k].cpu().clone()) for k, _ in model.named_parameters()]
    ref_opt = torch.optim.Adam(clones, lr=1e-3, weight_decay=0)
    ref_opt.load_state_dict(ckpt["optimizer"])
    for c, (k, _) in zip(clones, model.named_parameters()):
        c.grad = grads[k].cpu()
    ref_opt.step()
    lr = opt.param_groups[0]["lr"]
    for c, (k, p) in zip(clones, model.named_parameters()):
        assert float((p.detach().cpu() - c.detach()).abs().max()) <= 1e-3 * lr, k      # 1e-6 absolute: same arithmetic
    assert float(opt.state_dict()["state"][0]["step"]) == 3.0
    # (2) ... and lands where the REFERENCE's third step landed.  Adam divides by sqrt(v): where the gradient is rounding noise
    # (conv biases in front of a train-mode BatchNorm have a mathematically zero gradient) any two fp32 runs random-walk
    # apart by ~lr per step, so elements are compared where the loaded second moment says the gradient is real.
    state = ckpt["optimizer"]["state"]
    vmax = max(float(state[i]["exp_avg_sq"].max()) for i in range(len(state))) ** 0.5
    checked = 0
    for i, (k, p) in enumerate(model.named_parameters()):
        mask = state[i]["exp_avg_sq"].sqrt() > 1e-4 * vmax
        if not bool(mask.any()):
            continue
        d_ours = (p.detach().cpu() - before[k].cpu())[mask]
        d_ref = (torch.from_numpy(after["param:" + k]) - before[k].cpu())[mask]
        assert float((d_ours - d_ref).abs().max()) <= 0.05 * lr, k
        checked += int(mask.sum())
    assert checked > 10000


def test_fused_adam_state_round_trip(golden_dir):
    """FusedAdam -> state_dict -> torch.optim.Adam and back: both directions resume mid-run."""
    from titanet_b200.optim import FusedAdam
    ckpt, _ = _load(golden_dir)
    model, (x, y) = _model(ckpt)
    opt = FusedAdam(model.parameters(), lr=2e-3, betas=(0.8, 0.95), eps=1e-7, weight_decay=0.01)
    for _ in range(2):
        opt.zero_grad()
        model(x.cuda(), speakers=y.cuda())[2].backward()
        opt.step()
    sd = opt.state_dict()
    assert float(sd["state"][0]["step"]) == 2.0
    params = [p for p in model.parameters()]
    opt2 = FusedAdam(model.parameters(), lr=1e-3)
    opt2.load_state_dict(sd)
    opt2.param_groups[0]["betas"] = (0.7, 0.9)                   # hyper-parameters edited after load are honoured
    opt2.zero_grad()
    model(x.cuda(), speakers=y.cuda())[2].backward()            # (renormalises fc.weight in place: clone the parameters after it)
    torch_adam = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in params], lr=1e-3)
    torch_adam.load_state_dict(sd)                               # must not raise (KeyError 'step' before round 2)
    for q, p in zip(torch_adam.param_groups[0]["params"], params):
        q.grad = p.grad.detach().clone()
    torch_adam.param_groups[0]["betas"] = (0.7, 0.9)
    opt2.step()
    torch_adam.step()
    for q, p in zip(torch_adam.param_groups[0]["params"], params):
        assert float((p.detach() - q.detach()).abs().max()) <= 2e-6
