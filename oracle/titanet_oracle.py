"""CPU oracle for the TitaNet hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The shipped package
(``titanet_b200``) never does: its ops fail loudly when the CUDA library is
missing.

What it is: a *functional* restatement (plain ``torch`` CPU ops on a flat
``state_dict``; no ``nn.Module`` tree, no torchaudio) of the reference's
waveform -> mel -> encoder -> decoder -> loss path, written from the reference's
behaviour.  Every function cites the reference lines it follows
(paths relative to the reference checkout, Wadaboa/titanet @ 7b77053).

Parity pinning: the reference ships no tests / golden vectors, so
the oracle is pinned against *outputs of the reference itself*, produced in the
build container by ``oracle/make_golden.py`` (imports the reference modules from
``/root/reference/src``) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks this file against those vectors.

Third-party arithmetic restated here (absent from the reference tree): torchaudio
(pinned 0.13.0 in the reference's ``init/requirements.txt:4``; 2.11.0 in this
image) ``transforms.Spectrogram / MelScale / AmplitudeToDB`` and ``torch.stft``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, asdict
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# model description
# ----------------------------------------------------------------------------
@dataclass(frozen=True)
class TitaNetSpec:
    """Constructor arguments of ``models.TitaNet`` (src/models.py:175-192)."""

    n_mels: int = 80
    n_mega_blocks: int = 17
    n_sub_blocks: int = 3
    hidden: int = 256            # encoder_hidden_size
    enc_out: int = 1536          # encoder_output_size
    emb: int = 192               # embedding_size
    kernel: int = 3              # mega_block_kernel_size
    prolog_kernel: int = 3
    epilog_kernel: int = 1
    attn_hidden: int = 128
    se_reduction: int = 16
    simple_pool: bool = False
    dropout: float = 0.0

    @staticmethod
    def named(size: str, n_mega_blocks: int, **kw) -> "TitaNetSpec":
        """``TitaNet.get_titanet`` size table (src/models.py:310-316)."""
        h, k = {"s": (256, 3), "m": (512, 7), "l": (1024, 11)}[size.lower()]
        return TitaNetSpec(hidden=h, kernel=k, n_mega_blocks=n_mega_blocks, **kw)

    def to_dict(self):
        return asdict(self)


def _bn_keys(prefix: str, c: int):
    return [
        (prefix + ".weight", (c,)), (prefix + ".bias", (c,)),
        (prefix + ".running_mean", (c,)), (prefix + ".running_var", (c,)),
        (prefix + ".num_batches_tracked", ()),
    ]


def state_dict_schema(spec: TitaNetSpec, loss: Optional[str] = None,
                      n_classes: int = 0) -> "OrderedDict[str, tuple]":
    """Key -> shape map of ``TitaNet(...).state_dict()`` in registration order.

    Follows the module tree built in src/models.py:370-384 (encoder),
    435-455 (mega block), 497-513 (decoder), src/modules.py:64-79, 119-134,
    168-176 and src/losses.py:30, 70.  SURVEY.md §8(b) lists the schema.
    """
    H, D, A, E = spec.hidden, spec.enc_out, spec.attn_hidden, spec.emb
    ks: list = []
    p = "encoder.prolog.conv_block"
    ks += [(p + ".0.weight", (H, spec.n_mels, spec.prolog_kernel)), (p + ".0.bias", (H,))]
    ks += _bn_keys(p + ".1", H)
    for i in range(spec.n_mega_blocks):
        mb = f"encoder.mega_blocks.{i}"
        for j in range(spec.n_sub_blocks):
            q = f"{mb}.sub_blocks.{j}.conv_block"
            ks += [(q + ".0.conv.0.weight", (H, 1, spec.kernel)), (q + ".0.conv.0.bias", (H,)),
                   (q + ".0.conv.1.weight", (H, H, 1)), (q + ".0.conv.1.bias", (H,))]
            ks += _bn_keys(q + ".1", H)
        se = f"{mb}.sub_blocks.{spec.n_sub_blocks}.excitation"
        ks += [(se + ".0.weight", (H // spec.se_reduction, H)),
               (se + ".2.weight", (H, H // spec.se_reduction))]
        ks += [(mb + ".skip_connection.0.weight", (H, H, 1)), (mb + ".skip_connection.0.bias", (H,))]
        ks += _bn_keys(mb + ".skip_connection.1", H)
    p = "encoder.epilog.conv_block"
    ks += [(p + ".0.weight", (D, H, spec.epilog_kernel)), (p + ".0.bias", (D,))]
    ks += _bn_keys(p + ".1", D)
    if spec.simple_pool:
        ks += [("decoder.pool.2.weight", (2 * D, D)), ("decoder.pool.2.bias", (2 * D,))]
    else:
        ks += [("decoder.pool.0.in_linear.weight", (A, D)), ("decoder.pool.0.in_linear.bias", (A,)),
               ("decoder.pool.0.out_linear.weight", (D, A)), ("decoder.pool.0.out_linear.bias", (D,))]
        ks += _bn_keys("decoder.pool.1", 2 * D)
    ks += [("decoder.linear.0.weight", (E, 2 * D)), ("decoder.linear.0.bias", (E,))]
    ks += _bn_keys("decoder.linear.1", E)
    if loss is not None:
        ks += [("loss_function.fc.weight", (n_classes, E))]
        if loss == "ce":
            ks += [("loss_function.fc.bias", (n_classes,))]
    return OrderedDict(ks)


def synth_state_dict(spec: TitaNetSpec, loss: Optional[str] = None, n_classes: int = 0,
                     seed: int = 42, dtype=torch.float32) -> "OrderedDict[str, Tensor]":
    """Deterministic synthetic weights, independent of module construction order.

    Conv / linear tensors ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (the bound
    ``nn.Conv1d`` / ``nn.Linear`` default init uses); BatchNorm affine and running
    statistics are randomised (instead of 1/0/0/1) so eval-mode tests see
    non-trivial values.  One ``torch.Generator`` walks the schema in order.
    """
    g = torch.Generator().manual_seed(seed)
    schema = state_dict_schema(spec, loss, n_classes)
    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    for k, shape in schema.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.int64)
            continue
        prefix, leaf = k.rsplit(".", 1)
        is_bn = (prefix + ".running_mean") in schema
        u = torch.rand(shape, generator=g, dtype=torch.float64)
        if leaf == "running_var":
            v = 0.5 + u                                  # U(0.5, 1.5)
        elif leaf == "running_mean":
            v = 0.2 * (u - 0.5)                          # U(-0.1, 0.1)
        elif is_bn and leaf == "weight":
            v = 0.75 + 0.5 * u                           # gamma ~ U(0.75, 1.25)
        elif len(shape) == 1:                            # any bias (BN beta included)
            v = (2 * u - 1) / 4.0
        else:
            fan_in = 1
            for s_ in shape[1:]:
                fan_in *= s_
            v = (2 * u - 1) / math.sqrt(fan_in)
        sd[k] = v.to(dtype)
    return sd


# ----------------------------------------------------------------------------
# mel front end
# ----------------------------------------------------------------------------
def hann_window_periodic(win_length: int, dtype=torch.float32) -> Tensor:
    """``torch.hann_window(win_length)`` (periodic) as used by
    ``torchaudio.transforms.Spectrogram`` (src/transforms.py:134-140)."""
    n = torch.arange(win_length, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * n / win_length)).to(dtype)


def mel_filterbank(n_freqs: int = 257, n_mels: int = 80, sample_rate: int = 16000,
                   f_min: float = 0.0, f_max: Optional[float] = None) -> Tensor:
    """HTK triangular filterbank ``[n_freqs, n_mels]``, ``norm=None``.

    Restates ``torchaudio.functional.melscale_fbanks`` as reached through
    ``torchaudio.transforms.MelScale(n_mels, sample_rate, n_stft)``
    (src/transforms.py:142-144).  Computed in fp32 with the same operation order
    as torchaudio so the matrix is bit-identical.
    """
    if f_max is None:
        f_max = float(sample_rate // 2)
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.minimum(down, up), min=0.0)


def mel_spectrogram(waveform: Tensor, sample_rate: int = 16000, n_fft: int = 512,
                    win_length: int = 400, hop_length: int = 160, n_mels: int = 80) -> Tensor:
    """``transforms.MelSpectrogram.__call__`` without SpecAugment
    (src/transforms.py:158-184): waveform ``[C, L]`` -> ``[C, n_mels, 1 + L // hop]``.

    STFT: center=True with reflect padding of n_fft//2, periodic Hann(win_length)
    zero-padded symmetrically to n_fft, one-sided, un-normalised
    (``torchaudio.transforms.Spectrogram(power=None)``; line 165).  Then
    |.|^2 (178), HTK mel projection (182), 10*log10(clamp(.,1e-10)) (183),
    L2 normalisation over the mel axis with eps 1e-12 (184).
    """
    w = waveform if waveform.dim() == 2 else waveform.unsqueeze(0)
    pad = n_fft // 2
    x = F.pad(w.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)      # [C, L + n_fft]
    frames = x.unfold(-1, n_fft, hop_length)                               # [C, T, n_fft]
    win = torch.zeros(n_fft, dtype=w.dtype)
    off = (n_fft - win_length) // 2
    win[off:off + win_length] = hann_window_periodic(win_length, w.dtype)
    spec = torch.fft.rfft(frames * win, dim=-1)                            # [C, T, n_fft//2+1]
    power = spec.real ** 2 + spec.imag ** 2
    fb = mel_filterbank(n_fft // 2 + 1, n_mels, sample_rate).to(w.dtype)
    mel = power @ fb                                                       # [C, T, n_mels]
    db = 10.0 * torch.log10(torch.clamp(mel, min=1e-10))
    db = db.transpose(1, 2)                                                # [C, n_mels, T]
    return db / db.norm(p=2, dim=1, keepdim=True).clamp_min(1e-12)


def specaugment_draw(n_mels: int, n_frames: int, min_speed: float = 0.95, max_speed: float = 1.05,
                     freq_mask_ratio: float = 0.35, freq_mask_num: int = 1, time_mask_ratio: float = 0.15,
                     time_mask_num: int = 1, probability: float = 1.0):
    """The random draws of one ``MelSpectrogram.__call__`` with SpecAugment, in the reference's order and from the
    same generators (src/transforms.py:168-173: ``random.random()``, ``random.uniform``; 187-201 ->
    ``torchaudio.functional.mask_along_axis``: two ``torch.rand(1)`` per mask, ``value = rand * mask_param``,
    ``start = long(rand * (size - value))``, ``end = start + long(value)``; a mask_param < 1 draws nothing).
    Returns None (not applied) or ``(rate, stretched_frames, [(f0, f1), ...], [(t0, t1), ...])``."""
    import random
    if not (random.random() < probability):
        return None
    rate = random.uniform(min_speed, max_speed)
    frames = n_frames if rate == 1.0 else int(math.ceil(n_frames / rate))      # len(torch.arange(0, T, rate))

    def draw(mask_param, size):
        if mask_param < 1:
            return None
        value = torch.rand(1) * mask_param
        min_value = torch.rand(1) * (size - value)
        start = int(min_value.long())
        return start, start + int(value.long())

    fm = [m for m in (draw(freq_mask_ratio * n_mels, n_mels) for _ in range(freq_mask_num)) if m is not None]
    tm = [m for m in (draw(time_mask_ratio * frames, frames) for _ in range(time_mask_num)) if m is not None]
    return rate, frames, fm, tm


def mel_spectrogram_specaugment(waveform: Tensor, rate: float, freq_masks=(), time_masks=(), sample_rate: int = 16000,
                                n_fft: int = 512, win_length: int = 400, hop_length: int = 160, n_mels: int = 80) -> Tensor:
    """``transforms.MelSpectrogram.__call__`` WITH SpecAugment for given draws (src/transforms.py:165-201).

    ``torchaudio.transforms.TimeStretch`` (phase vocoder, torchaudio 0.13 ``functional.phase_vocoder``) resamples the
    complex STFT at ``time_steps = arange(0, T, rate)``: magnitude ``alpha |X[i+1]| + (1 - alpha) |X[i]|`` with
    ``alpha = time_steps % 1`` and two zero frames appended, phase accumulated separately.  Line 178 takes
    ``abs().pow(2)`` right after, so the phase never reaches the mel spectrogram: only the interpolated magnitude does.
    Then mel / dB / L2-normalise as without augmentation, then the masks set rows [f0, f1) and frames [t0, t1) to 0."""
    w = waveform if waveform.dim() == 2 else waveform.unsqueeze(0)
    pad = n_fft // 2
    x = F.pad(w.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    frames = x.unfold(-1, n_fft, hop_length)
    win = torch.zeros(n_fft, dtype=w.dtype)
    off = (n_fft - win_length) // 2
    win[off:off + win_length] = hann_window_periodic(win_length, w.dtype)
    spec = torch.fft.rfft(frames * win, dim=-1)                            # [C, T, F]
    mag = spec.abs()
    if rate != 1.0:
        T = mag.shape[1]
        steps = torch.arange(0, T, rate, dtype=w.dtype)
        alphas = (steps % 1.0).view(1, -1, 1)
        mag = F.pad(mag, (0, 0, 0, 2))
        mag = alphas * mag.index_select(1, (steps + 1).long()) + (1 - alphas) * mag.index_select(1, steps.long())
    power = mag ** 2
    fb = mel_filterbank(n_fft // 2 + 1, n_mels, sample_rate).to(w.dtype)
    db = 10.0 * torch.log10(torch.clamp(power @ fb, min=1e-10)).transpose(1, 2)
    out = db / db.norm(p=2, dim=1, keepdim=True).clamp_min(1e-12)
    for f0, f1 in freq_masks:
        out[:, f0:f1, :] = 0.0
    for t0, t1 in time_masks:
        out[:, :, t0:t1] = 0.0
    return out


def collate_pad(mels) -> Tuple[Tensor, Tensor]:
    """``datasets.collate_fn`` (src/datasets.py:48-73): zero-pad ``[1, M, T_i]`` mels
    to the batch max T; returns (``[B, M, Tmax]``, lengths)."""
    tmax = max(m.shape[-1] for m in mels)
    out = torch.zeros(len(mels), mels[0].shape[-2], tmax, dtype=mels[0].dtype)
    for i, m in enumerate(mels):
        out[i, :, : m.shape[-1]] = m.reshape(m.shape[-2], m.shape[-1])
    return out, torch.tensor([m.shape[-1] for m in mels])


# ----------------------------------------------------------------------------
# encoder / decoder
# ----------------------------------------------------------------------------
class _Ctx:
    """Bookkeeping for one functional forward: mode, dropout, running-stat updates."""

    def __init__(self, sd, training, dropout, momentum=0.1, update_running=True):
        self.sd, self.training, self.p = sd, training, dropout
        self.momentum, self.update_running = momentum, update_running
        self.new_running: Dict[str, Tensor] = {}

    def bn(self, x: Tensor, prefix: str, eps: float = 1e-5) -> Tensor:
        """``nn.BatchNorm1d`` (src/modules.py:128, src/models.py:454,506,512): batch
        statistics (biased var) in training mode, running statistics in eval mode;
        running stats move with momentum 0.1 using the unbiased variance."""
        sd = self.sd
        g, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
        rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
        dims = (0,) if x.dim() == 2 else (0, 2)
        shape = (1, -1) if x.dim() == 2 else (1, -1, 1)
        if self.training:
            n = x.numel() // x.shape[1]
            mean = x.mean(dim=dims)
            var = x.var(dim=dims, unbiased=False)
            if self.update_running:
                with torch.no_grad():
                    m = self.momentum
                    self.new_running[prefix + ".running_mean"] = (1 - m) * rm + m * mean.detach()
                    self.new_running[prefix + ".running_var"] = (1 - m) * rv + m * var.detach() * n / max(n - 1, 1)
                    self.new_running[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
        else:
            mean, var = rm, rv
        return (x - mean.view(shape)) / torch.sqrt(var.view(shape) + eps) * g.view(shape) + b.view(shape)

    def drop(self, x: Tensor) -> Tensor:
        return F.dropout(x, p=self.p, training=self.training) if self.p > 0 else x


def conv1d_same(x: Tensor, w: Tensor, b: Optional[Tensor], groups: int = 1) -> Tensor:
    """``Conv1dSamePadding.forward`` (src/modules.py:14-40) for stride 1 / dilation 1:
    zero-pad (K-1)//2 on both sides, then a padding-free conv."""
    k = w.shape[-1]
    pad = (k - 1) // 2          # (S*(W-1) - W + K + (D-1)(K-1)) // 2 with S = D = 1
    return F.conv1d(F.pad(x, (pad, pad)), w, b, groups=groups)


def encoder_forward(ctx: _Ctx, spec: TitaNetSpec, x: Tensor) -> Tensor:
    """``Encoder.forward`` (src/models.py:386-404) / ``MegaBlock.forward`` (457-472)."""
    sd = ctx.sd
    p = "encoder.prolog.conv_block"
    # prolog: conv -> BN -> ReLU, no dropout (ConvBlock1d default dropout=0; models.py:370)
    x = torch.relu(ctx.bn(conv1d_same(x, sd[p + ".0.weight"], sd[p + ".0.bias"]), p + ".1"))
    for i in range(spec.n_mega_blocks):
        mb = f"encoder.mega_blocks.{i}"
        # skip branch: plain nn.Conv1d(k=1) -> BN (models.py:452-455); evaluated first (468)
        skip = ctx.bn(F.conv1d(x, sd[mb + ".skip_connection.0.weight"], sd[mb + ".skip_connection.0.bias"]),
                      mb + ".skip_connection.1")
        y = x
        for j in range(spec.n_sub_blocks):
            q = f"{mb}.sub_blocks.{j}.conv_block"
            # DepthwiseConv1d: depthwise(K, groups=C, bias) then pointwise 1x1 (modules.py:64-79)
            y = conv1d_same(y, sd[q + ".0.conv.0.weight"], sd[q + ".0.conv.0.bias"], groups=y.shape[1])
            y = conv1d_same(y, sd[q + ".0.conv.1.weight"], sd[q + ".0.conv.1.bias"])
            y = ctx.drop(torch.relu(ctx.bn(y, q + ".1")))          # BN -> ReLU -> Dropout (modules.py:128-133)
        se = f"{mb}.sub_blocks.{spec.n_sub_blocks}.excitation"
        # SqueezeExcitation (modules.py:173-189): mean_T -> Linear -> ReLU -> Linear -> Sigmoid -> scale
        g = torch.sigmoid(torch.relu(y.mean(dim=2) @ sd[se + ".0.weight"].t()) @ sd[se + ".2.weight"].t())
        y = y * g.unsqueeze(-1)
        x = ctx.drop(torch.relu(skip + y))                         # models.py:468-472
    p = "encoder.epilog.conv_block"
    return torch.relu(ctx.bn(conv1d_same(x, sd[p + ".0.weight"], sd[p + ".0.bias"]), p + ".1"))


def attentive_stats_pooling(sd, x: Tensor, eps: float = 1e-6) -> Tensor:
    """``AttentiveStatsPooling.forward`` (src/models.py:553-584)."""
    wi, bi = sd["decoder.pool.0.in_linear.weight"], sd["decoder.pool.0.in_linear.bias"]
    wo, bo = sd["decoder.pool.0.out_linear.weight"], sd["decoder.pool.0.out_linear.bias"]
    xt = x.transpose(1, 2)                                         # [B, T, DE]
    e = (torch.tanh(xt @ wi.t() + bi) @ wo.t() + bo).transpose(1, 2)
    a = torch.softmax(e, dim=2)
    mu = (a * x).sum(dim=2)
    resid = (a * x ** 2).sum(dim=2) - mu ** 2
    return torch.cat([mu, torch.sqrt(resid.clamp(min=eps))], dim=1)


def decoder_forward(ctx: _Ctx, spec: TitaNetSpec, enc: Tensor) -> Tensor:
    """``Decoder.forward`` (src/models.py:515-529; construction 495-513)."""
    sd = ctx.sd
    if spec.simple_pool:
        pooled = enc.mean(dim=2) @ sd["decoder.pool.2.weight"].t() + sd["decoder.pool.2.bias"]
    else:
        pooled = ctx.bn(attentive_stats_pooling(sd, enc), "decoder.pool.1")
    lin = pooled @ sd["decoder.linear.0.weight"].t() + sd["decoder.linear.0.bias"]
    return ctx.bn(lin, "decoder.linear.1")


# ----------------------------------------------------------------------------
# loss heads
# ----------------------------------------------------------------------------
def ce_head(sd, emb: Tensor, targets: Tensor):
    """``CELoss.forward`` (src/losses.py:32-44)."""
    logits = emb @ sd["loss_function.fc.weight"].t() + sd["loss_function.fc.bias"]
    preds = torch.argmax(logits, dim=1)
    loss = F.cross_entropy(logits, targets)
    return F.normalize(emb, p=2, dim=1), preds, loss


def angular_margin_head(sd, emb: Tensor, targets: Tensor, scale: Optional[float],
                        m1: float = 1.0, m2: float = 0.0, m3: float = 0.0, eps: float = 1e-6):
    """``AngularMarginLoss.forward`` (src/losses.py:77-132) in closed form.

    Quirks kept: the class weights are row-normalised in place, outside autograd
    (line 86; the caller sees ``sd['loss_function.fc.weight']`` overwritten); the
    input norm has no eps (89-92); ``scale=None`` uses the per-sample norm (95-99);
    no max-subtraction before ``exp`` (127); ``eps`` is added to the denominator (130).
    The per-row python ``torch.cat`` loop (119-126) is the sum over j != y_i.
    """
    w = sd["loss_function.fc.weight"]
    with torch.no_grad():
        w.copy_(F.normalize(w.detach(), p=2, dim=1))
    norms = torch.norm(emb, p=2, dim=1)
    xh = emb / norms.unsqueeze(-1)
    scales = torch.full_like(norms, float(scale)) if scale is not None else norms
    cos = (xh @ w.t()).clamp(-1, 1)
    preds = torch.argmax(cos, dim=1)
    cos_y = cos.gather(1, targets.view(-1, 1)).squeeze(1)
    num = scales * (torch.cos(m1 * torch.arccos(cos_y) + m2) - m3)
    ex = torch.exp(scales.unsqueeze(-1) * cos)
    # sum over j != y_i (explicit exclusion, like the reference's torch.cat of both sides)
    mask = torch.ones_like(cos, dtype=torch.bool)
    mask.scatter_(1, targets.view(-1, 1), False)
    others = (ex * mask).sum(dim=1)
    den = torch.exp(num) + others
    loss = -torch.mean(num - torch.log(den + eps))
    return xh, preds, loss


LOSS_KINDS = ("ce", "arc", "cos", "sphere")


def loss_margins(kind: str, margin: float) -> Tuple[float, float, float]:
    """(m1, m2, m3) per head: SphereFace (losses.py:135-149), CosFace (152-166),
    ArcFace (169-183)."""
    return {"sphere": (margin, 0.0, 0.0), "arc": (1.0, margin, 0.0), "cos": (1.0, 0.0, margin)}[kind]


# ----------------------------------------------------------------------------
# whole model
# ----------------------------------------------------------------------------
def titanet_forward(sd, spec: TitaNetSpec, spectrograms: Tensor, speakers: Optional[Tensor] = None,
                    loss: Optional[str] = None, scale: Optional[float] = None, margin: float = 0.0,
                    training: bool = False, update_running: bool = True):
    """``TitaNet.forward`` (src/models.py:318-339).

    Returns ``(embeddings,)`` (unit norm) when ``speakers`` is None, else
    ``(embeddings, preds, loss)``; second return value is the dict of updated
    BatchNorm buffers (training mode) that the caller may merge into ``sd``.
    """
    ctx = _Ctx(sd, training, spec.dropout, update_running=update_running)
    enc = encoder_forward(ctx, spec, spectrograms)
    emb = decoder_forward(ctx, spec, enc)
    if speakers is None:
        return (F.normalize(emb, p=2, dim=1),), ctx.new_running
    assert loss is not None, "Loss function should not be None in training mode"
    if loss == "ce":
        out = ce_head(sd, emb, speakers)
    else:
        m1, m2, m3 = loss_margins(loss, margin)
        out = angular_margin_head(sd, emb, speakers, scale, m1, m2, m3)
    return out, ctx.new_running


def titanet_step(sd, spec: TitaNetSpec, spectrograms: Tensor, speakers: Tensor, loss: str,
                 scale: Optional[float] = None, margin: float = 0.0, training: bool = True,
                 input_grad: bool = False):
    """fwd + bwd (the ``learn.py:95-117`` step without the optimizer): returns
    ``(emb, preds, loss, grads{name: tensor}[, dx])``.  Uses autograd on the
    functional forward; works in fp32 or fp64 depending on ``sd`` dtype."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
    work = OrderedDict((k, params.get(k, v)) for k, v in sd.items())
    if loss != "ce":            # in-place renorm happens on .data outside autograd (losses.py:86)
        with torch.no_grad():
            w = params["loss_function.fc.weight"]
            w.copy_(F.normalize(w, p=2, dim=1))
            sd["loss_function.fc.weight"].copy_(w)
    x = spectrograms.detach().clone().requires_grad_(input_grad)
    (emb, preds, lval), new_running = titanet_forward(work, spec, x, speakers, loss, scale, margin,
                                                      training=training)
    lval.backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}
    out = (emb.detach(), preds, lval.detach(), grads, new_running)
    return out + ((x.grad,) if input_grad else ())


def synthetic_batch(batch: int, seconds: float = 3.0, n_classes: int = 251, seed: int = 42,
                    sample_rate: int = 16000):
    """Synthetic inputs of SURVEY.md §8(d): waveform 0.1*randn(B, L), labels randint(0, C)."""
    g = torch.Generator().manual_seed(seed)
    wave = 0.1 * torch.randn(batch, int(round(seconds * sample_rate)), generator=g)
    labels = torch.randint(0, n_classes, (batch,), generator=g)
    return wave, labels
