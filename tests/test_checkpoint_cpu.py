"""Checkpoint format compatibility: the reference saves
``{"model": state_dict, "optimizer": ..., "lr_scheduler": ..., "epoch": n}`` (src/learn.py:180-201) and the notebook
loads ``state_dict["model"]`` with ``strict=False`` into a model built WITHOUT the loss head, which reports
``unexpected_keys=['loss_function.fc.weight']`` (titanet.ipynb:1379-1382, output at :1370).  The drop-in modules must
round-trip that file format; FusedAdam must accept a torch.optim.Adam state dict (resume)."""
import io

import torch

import titanet_oracle as O
from titanet_b200 import losses, models
from titanet_b200.optim import FusedAdam


def _model(loss=True, blocks=2):
    lf = losses.ArcFaceLoss(192, 251, scale=30, margin=0.2) if loss else None
    return models.TitaNet.get_titanet(192, 80, blocks, "s", loss_function=lf, dropout=0.1)


def test_reference_checkpoint_dict_round_trips():
    spec = O.TitaNetSpec.named("s", 2)
    sd = O.synth_state_dict(spec, "arc", 251)                    # reference key schema (pinned by oracle/make_golden.py)
    src = _model()
    src.load_state_dict(sd, strict=True)
    opt = torch.optim.Adam(src.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10)
    buf = io.BytesIO()
    torch.save({"model": src.state_dict(), "optimizer": opt.state_dict(), "lr_scheduler": sched.state_dict(), "epoch": 25}, buf)
    buf.seek(0)
    ckpt = torch.load(buf, map_location="cpu")
    assert set(ckpt) == {"model", "optimizer", "lr_scheduler", "epoch"} and ckpt["epoch"] == 25
    dst = _model()
    res = dst.load_state_dict(ckpt["model"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in dst.state_dict().items():
        assert torch.equal(v, sd[k]), k
    # the notebook's inference load: no loss head on the receiving model
    infer = _model(loss=False)
    res = infer.load_state_dict(ckpt["model"], strict=False)
    assert res.missing_keys == [] and res.unexpected_keys == ["loss_function.fc.weight"]


def test_fused_adam_accepts_torch_adam_state_dict():
    net = _model(blocks=1)
    ref = torch.optim.Adam(net.parameters(), lr=2e-3, betas=(0.8, 0.95), eps=1e-7, weight_decay=0.01)
    ours = FusedAdam(net.parameters(), lr=1e-3)
    ours.load_state_dict(ref.state_dict())
    g = ours.param_groups[0]
    assert (g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]) == (2e-3, (0.8, 0.95), 1e-7, 0.01)
    assert len(g["params"]) == len(list(net.parameters()))
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(ours, T_max=4)      # train.py:138-144
    for _ in range(2):
        sched.step()
    assert ours.param_groups[0]["lr"] < 2e-3
