"""Drop-in for the mel front end of the reference's ``src/transforms.py``.

``MelSpectrogram`` keeps the reference's constructor signature and its
example-dict call contract (src/transforms.py:111-203) and adds a batched entry point
(``batch``) that turns ``[B, L]`` waveforms (+ per-utterance lengths) into the padded
``[B, n_mels, T]`` tensor ``datasets.collate_fn`` would build (src/datasets.py:48-73).
The arithmetic is one CUDA kernel (csrc/mel.cu).  The window and the HTK filterbank
are small host-side constants computed here with the formulas torchaudio uses
(``torch.hann_window``; ``torchaudio.functional.melscale_fbanks``, htk, norm=None).

SpecAugment (src/transforms.py:168-175, 187-201) runs inside the same
kernel: the random draws (apply?, stretch rate, mask ranges) are made on the host from the
generators the reference uses (``random`` and ``torch.rand``) in the reference's order, so
equal seeds give equal augmentations; the kernel receives them as per-utterance arrays.

Out of the accelerated path (DESIGN.md "scope"): ``SpeedPerturbation``, ``Reverb``;
``Resample`` only accepts inputs already at the target rate.
"""
from __future__ import annotations

import math
import random

import torch

from . import _ops as ops
from ._lib import TitanetLibraryError


def copy_example(example):
    """Copy a dataset example, cloning its tensors (reference: src/transforms.py:12-22)."""
    return {k: (torch.clone(v) if isinstance(v, torch.Tensor) else v) for k, v in example.items()}


class SpecAugmentDraw:
    """The random draws of one augmented utterance: stretch ``rate``, stretched frame count, frequency masks and
    time masks as half-open ``(start, end)`` ranges."""

    __slots__ = ("rate", "frames", "freq_masks", "time_masks")

    def __init__(self, rate: float, frames: int, freq_masks=(), time_masks=()):
        self.rate, self.frames = float(rate), int(frames)
        self.freq_masks, self.time_masks = list(freq_masks), list(time_masks)

    def __repr__(self):
        return f"SpecAugmentDraw(rate={self.rate}, frames={self.frames}, freq={self.freq_masks}, time={self.time_masks})"


def stretched_frames(n_frames: int, rate: float) -> int:
    """Frames after ``torchaudio.transforms.TimeStretch``: ``len(torch.arange(0, n_frames, rate))`` (bypass at rate 1)."""
    return n_frames if rate == 1.0 else int(math.ceil(n_frames / rate))


def _draw_mask(mask_param: float, size: int):
    """``torchaudio.functional.mask_along_axis``' draws (two ``torch.rand(1)``); nothing is drawn when mask_param < 1."""
    if mask_param < 1:
        return None
    value = torch.rand(1) * mask_param
    min_value = torch.rand(1) * (size - value)
    start = int(min_value.long())
    return start, start + int(value.long())


def _htk_filterbank(n_freqs: int, n_mels: int, sample_rate: int) -> torch.Tensor:
    """[n_freqs, n_mels] triangular HTK mel filters over 0 .. sample_rate/2, no area norm."""
    freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    mel_max = 2595.0 * math.log10(1.0 + float(sample_rate // 2) / 700.0)
    mel_pts = torch.linspace(0.0, mel_max, n_mels + 2)
    hz_pts = 700.0 * (10.0 ** (mel_pts / 2595.0) - 1.0)
    width = hz_pts[1:] - hz_pts[:-1]
    dist = hz_pts.unsqueeze(0) - freqs.unsqueeze(1)
    rising = (-1.0 * dist[:, :-2]) / width[:-1]
    falling = dist[:, 2:] / width[1:]
    return torch.clamp(torch.minimum(rising, falling), min=0.0)


class MelSpectrogram:
    """Waveform -> L2-normalised log-mel spectrogram (reference: src/transforms.py:111-203)."""

    def __init__(self, sample_rate, n_fft=400, win_length=None, hop_length=None, n_mels=128, specaugment_min_speed=0.95,
                 specaugment_max_speed=1.05, specaugment_freq_mask_ratio=0.35, specaugment_freq_mask_num=1,
                 specaugment_time_mask_ratio=0.15, specaugment_time_mask_num=1, specaugment_probability=1.0):
        self.sample_rate = sample_rate
        self.n_fft = n_fft
        self.win_length = win_length if win_length is not None else n_fft          # torchaudio defaults
        self.hop_length = hop_length if hop_length is not None else self.win_length // 2
        self.n_mels = n_mels
        self.specaugment_min_speed = specaugment_min_speed
        self.specaugment_max_speed = specaugment_max_speed
        self.specaugment_freq_mask_ratio = specaugment_freq_mask_ratio
        self.specaugment_freq_mask_num = specaugment_freq_mask_num
        self.specaugment_time_mask_ratio = specaugment_time_mask_ratio
        self.specaugment_time_mask_num = specaugment_time_mask_num
        self.specaugment_probability = specaugment_probability
        if self.n_fft & (self.n_fft - 1) or not 64 <= self.n_fft <= 4096:
            raise NotImplementedError(f"n_fft={n_fft}: the CUDA STFT needs a power of two in [64, 4096] "
                                      "(the reference's training configuration uses 512)")
        if self.win_length > self.n_fft:
            raise ValueError("win_length must be <= n_fft")
        # host-side constants
        window = torch.zeros(self.n_fft)
        off = (self.n_fft - self.win_length) // 2
        window[off:off + self.win_length] = torch.hann_window(self.win_length, periodic=True)
        fb = _htk_filterbank(self.n_fft // 2 + 1, n_mels, sample_rate)
        nz = fb > 0
        lo = torch.where(nz.any(0), nz.float().argmax(0), torch.zeros(n_mels, dtype=torch.long))
        hi = torch.where(nz.any(0), fb.shape[0] - nz.flip(0).float().argmax(0), torch.zeros(n_mels, dtype=torch.long))
        self._host = (window, fb.contiguous(), lo.to(torch.int32), hi.to(torch.int32))
        self._dev = {}

    def _consts(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = tuple(t.to(device) for t in self._host)
        return self._dev[key]

    def n_frames(self, n_samples: int) -> int:
        return 1 + n_samples // self.hop_length

    def draw_specaugment(self, n_frames: int):
        """One utterance's SpecAugment draws, from the generators and in the order the reference consumes them
        (src/transforms.py:168-173 ``random.random()`` / ``random.uniform``; 187-201 ``torch.rand`` inside
        ``mask_along_axis``).  Returns ``None`` when the coin flip says "no augmentation"."""
        if not (random.random() < self.specaugment_probability):
            return None
        rate = random.uniform(self.specaugment_min_speed, self.specaugment_max_speed)
        frames = stretched_frames(n_frames, rate)
        fm = [_draw_mask(self.specaugment_freq_mask_ratio * self.n_mels, self.n_mels)
              for _ in range(self.specaugment_freq_mask_num)]
        tm = [_draw_mask(self.specaugment_time_mask_ratio * frames, frames) for _ in range(self.specaugment_time_mask_num)]
        return SpecAugmentDraw(rate, frames, [m for m in fm if m], [m for m in tm if m])

    def batch(self, waveforms: torch.Tensor, lengths: torch.Tensor | None = None, channels_last: bool = False,
              augment=None) -> torch.Tensor:
        """``[B, L]`` CUDA waveforms -> ``[B, n_mels, T]`` with ``T = 1 + L // hop``.  With
        ``lengths`` (samples per utterance) each utterance is transformed on its own length
        (own reflect padding) and frames past it are zero, like ``collate_fn``.  ``lengths`` is device data (a static
        buffer of a captured step), so it is validated on the device: a length beyond ``L`` is clipped to ``L`` and a
        length <= ``n_fft // 2`` (which ``torch.stft`` rejects) produces an all-zero spectrogram.

        ``augment``: ``None`` (no SpecAugment), ``True`` (draw per utterance with ``draw_specaugment``) or a list of
        B ``SpecAugmentDraw`` / ``None``.  Stretched utterances have ``ceil(T_b / rate_b)`` frames; T becomes the
        batch maximum and shorter utterances are zero padded, again like ``collate_fn``."""
        if waveforms.dim() != 2:
            raise ValueError("expected [B, L] waveforms")
        if lengths is None and waveforms.shape[1] <= self.n_fft // 2:
            raise ValueError("waveform shorter than n_fft // 2 + 1 samples (reflect padding)")
        window, fb, lo, hi = self._consts(waveforms.device)
        if augment is None:
            return ops.mel_forward(waveforms, lengths, window, fb, lo, hi, self.n_fft, self.hop_length, self.n_mels,
                                   nwc=channels_last)
        B, L = waveforms.shape
        lens = [L] * B if lengths is None else [int(v) for v in lengths.tolist()]
        if any(v <= self.n_fft // 2 or v > L for v in lens):      # host-visible here: refuse what torch.stft refuses
            raise ValueError(f"utterance lengths must lie in ({self.n_fft // 2}, {L}] samples (reflect padding / row length)")
        src_frames = [self.n_frames(v) for v in lens]
        if augment is True:
            augment = [self.draw_specaugment(t) for t in src_frames]
        if len(augment) != B:
            raise ValueError("augment must hold one entry per utterance")
        nf = max([len(d.freq_masks) for d in augment if d is not None] + [0])
        nt = max([len(d.time_masks) for d in augment if d is not None] + [0])
        rates = torch.ones(B, dtype=torch.float64)
        frames = torch.tensor(src_frames, dtype=torch.int32)
        masks = torch.zeros(B, max(nf + nt, 1), 2, dtype=torch.int32)
        for b, d in enumerate(augment):
            if d is None:
                continue
            if d.frames != stretched_frames(src_frames[b], d.rate):
                raise ValueError(f"utterance {b}: draw made for another length ({d.frames} frames)")
            rates[b], frames[b] = d.rate, d.frames
            for q, m in enumerate(d.freq_masks):
                masks[b, q] = torch.tensor(m, dtype=torch.int32)
            for q, m in enumerate(d.time_masks):
                masks[b, nf + q] = torch.tensor(m, dtype=torch.int32)
        dev = waveforms.device
        return ops.mel_forward(waveforms, lengths, window, fb, lo, hi, self.n_fft, self.hop_length, self.n_mels,
                               T_out=int(frames.max()), nwc=channels_last, rates=rates.to(dev), frames=frames.to(dev),
                               masks=masks.to(dev) if nf + nt else None, n_fmask=nf, n_tmask=nt)

    def __call__(self, example):
        assert isinstance(example, dict) and "waveform" in example, "Wrong input structure"
        new_example = copy_example(example)
        wave = new_example["waveform"]
        src_device = wave.device
        if wave.dim() == 1:
            wave = wave.unsqueeze(0)
        if src_device.type != "cuda" and not torch.cuda.is_available():
            raise TitanetLibraryError("titanet_b200.transforms.MelSpectrogram runs on NVIDIA B200 (sm_100a) only: no CUDA "
                                      "device is visible and there is no CPU fallback")
        dev = src_device if src_device.type == "cuda" else torch.device("cuda")
        draw = self.draw_specaugment(self.n_frames(wave.shape[-1]))                 # one draw per example, like the reference
        spec = self.batch(wave.to(device=dev, dtype=torch.float32),
                          augment=None if draw is None else [draw] * wave.shape[0])  # [C, n_mels, T]
        new_example["spectrogram"] = spec.to(src_device)
        return new_example


class RandomChunk:
    """Random crop of utterances longer than ``max_length`` seconds (reference:
    src/transforms.py:206-233); slicing only, no kernel."""

    def __init__(self, max_length, lengths):
        self.max_length = max_length
        self.lengths = lengths

    def __call__(self, example):
        assert isinstance(example, dict) and "waveform" in example and "sample_rate" in example, "Wrong input structure"
        new_example = copy_example(example)
        num_samples = new_example["waveform"].size(-1)
        if num_samples / new_example["sample_rate"] > self.max_length:
            length = random.choice(self.lengths)
            samples = int(length * new_example["sample_rate"])
            start = random.randint(0, num_samples - samples)
            new_example["waveform"] = new_example["waveform"][:, start:start + samples]
        return new_example


class Resample:
    """Identity for inputs already at the target rate; anything else is out of scope."""

    def __init__(self, target_sample_rate):
        self.target_sample_rate = target_sample_rate

    def __call__(self, example):
        assert isinstance(example, dict) and "waveform" in example and "sample_rate" in example, "Wrong input structure"
        if example["sample_rate"] != self.target_sample_rate:
            raise NotImplementedError("resampling is outside the titanet_b200 hot path; feed 16 kHz audio")
        return copy_example(example)


def get_transforms(enabled, rir_corpora_path, max_length=3, chunk_lengths=[1.5, 2, 3], min_speed=0.95, max_speed=1.05,
                   sample_rate=16000, n_fft=512, win_length=25, hop_length=10, n_mels=80, freq_mask_ratio=0.35,
                   freq_mask_num=1, time_mask_ratio=0.15, time_mask_num=1, probability=1.0, device="cpu", training=True):
    """Transformation list of the TitaNet paper (reference: src/transforms.py:25-75)."""
    if enabled is None:
        enabled = []
    transformations = [Resample(sample_rate)]
    if "chunk" in enabled:
        transformations += [RandomChunk(max_length, chunk_lengths)]
    if "reverb" in enabled and training:
        raise NotImplementedError("Reverb augmentation is outside the titanet_b200 hot path")
    transformations += [
        MelSpectrogram(sample_rate, n_fft=n_fft, win_length=int(win_length / 1000 * sample_rate),
                       hop_length=int(hop_length / 1000 * sample_rate), n_mels=n_mels, specaugment_min_speed=min_speed,
                       specaugment_max_speed=max_speed, specaugment_freq_mask_ratio=freq_mask_ratio,
                       specaugment_freq_mask_num=freq_mask_num, specaugment_time_mask_ratio=time_mask_ratio,
                       specaugment_time_mask_num=time_mask_num,
                       specaugment_probability=(probability if "specaugment" in enabled and training else 0.0))
    ]
    return transformations
