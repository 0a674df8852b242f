"""Seeded synthetic cases shared by ``oracle/make_golden.py`` and ``tests/``.
TEST INFRASTRUCTURE (see titanet_oracle.py header)."""
import torch

from titanet_oracle import TitaNetSpec

TINY = dict(
    tiny_k3=TitaNetSpec(hidden=64, kernel=3, n_mega_blocks=2, enc_out=96, attn_hidden=32, emb=48),
    tiny_k7=TitaNetSpec(hidden=32, kernel=7, n_mega_blocks=1, enc_out=64, attn_hidden=16, emb=32),
    tiny_k11=TitaNetSpec(hidden=32, kernel=11, n_mega_blocks=1, enc_out=64, attn_hidden=16, emb=32),
    # Decoder(simple_pool=True): average pooling + Linear instead of attentive statistics (models.py:497-502)
    tiny_k3_simple=TitaNetSpec(hidden=64, kernel=3, n_mega_blocks=1, enc_out=96, attn_hidden=32, emb=48, simple_pool=True),
)

# name -> (spec, loss, n_classes, B, T, scale, margin, full_grads)
TRAIN_CASES = {
    "tiny_k3_ce": (TINY["tiny_k3"], "ce", 10, 4, 50, None, 0.0, True),
    "tiny_k3_arc": (TINY["tiny_k3"], "arc", 10, 4, 50, 30, 0.2, True),
    "tiny_k3_cos": (TINY["tiny_k3"], "cos", 10, 4, 37, 64, 0.2, True),
    "tiny_k3_arc_noscale": (TINY["tiny_k3"], "arc", 10, 4, 37, None, 0.2, True),
    "tiny_k7_ce": (TINY["tiny_k7"], "ce", 7, 3, 33, None, 0.0, True),
    "tiny_k11_arc": (TINY["tiny_k11"], "arc", 7, 3, 64, 30, 0.2, True),
    "tiny_k3_simple_ce": (TINY["tiny_k3_simple"], "ce", 10, 4, 41, None, 0.0, True),
    "s17_ce_b4": (TitaNetSpec.named("s", 17), "ce", 251, 4, 101, None, 0.0, False),
}

# pinned for the oracle only (CPU): the GPU suite parametrises over TRAIN_CASES
ORACLE_ONLY_CASES = {
    # SphereFace: multiplicative angular margin cos(m1 * theta), default margin 3 (src/losses.py:135-149)
    "tiny_k3_sphere": (TINY["tiny_k3"], "sphere", 10, 4, 43, 30, 3, True),
    # eval-free train step at the notebook's class count with the loss's default scale (64) and margin (0.5)
    "tiny_k7_arc_s64": (TINY["tiny_k7"], "arc", 251, 5, 29, 64, 0.5, True),
}


# BASELINE.json configs[2] / [3] model families at batches the CPU reference finishes in seconds: TitaNet-M/10 (hidden 512,
# depthwise K = 7) and TitaNet-L/5 (hidden 1024, K = 11) with the ArcFace head of parameters.yml:42-44 (s = 30, m = 0.2); the
# L case is "ragged": every utterance has its own number of frames (1..8 s) and zeros behind it, as datasets.collate_fn
# (src/datasets.py:48-73) leaves a padded batch -- the model never sees the lengths (src/learn.py:88).
BIG_CASES = {
    "m10_arc_b8": dict(spec=TitaNetSpec.named("m", 10), loss="arc", nc=251, B=8, T=301, scale=30, margin=0.2, ragged=False),
    "l5_arc_ragged_b4": dict(spec=TitaNetSpec.named("l", 5), loss="arc", nc=251, B=4, T=801, scale=30, margin=0.2, ragged=True),
}


def big_inputs(case, seed=42):
    """(x [B, 80, T], labels, frames per utterance) of a BIG_CASES entry."""
    x, y = train_inputs(case["spec"], case["nc"], case["B"], case["T"], seed)
    frames = torch.full((case["B"],), case["T"], dtype=torch.int64)
    if case["ragged"]:
        g = torch.Generator().manual_seed(seed + 3)
        frames = 1 + 100 * torch.randint(1, 9, (case["B"],), generator=g)
        frames[0] = case["T"]                                  # the batch maximum defines T
        for b in range(case["B"]):
            x[b, :, int(frames[b]):] = 0.0
    return x, y, frames


def train_inputs(spec, n_classes, B, T, seed=42):
    g = torch.Generator().manual_seed(seed + 1)
    x = 0.3 * torch.randn(B, spec.n_mels, T, generator=g)
    y = torch.randint(0, n_classes, (B,), generator=g)
    return x, y


def mel_inputs():
    g = torch.Generator().manual_seed(7)
    odd = 0.1 * torch.randn(12345, generator=g)
    silence = torch.zeros(4000)
    loud = 3.0 * torch.randn(8000, generator=g)
    return dict(mel_odd=odd, mel_silence=silence, mel_loud=loud)


def eval_dx_inputs():
    g = torch.Generator().manual_seed(5)
    return torch.randn(3, 80, 40, generator=g)


# SpecAugment cases: (samples, seed for ``random`` and ``torch``); the waveform comes from ``specaug_wave``
SPECAUG_CASES = [(16000, 101), (48000, 102), (12345, 103), (30001, 104), (20000, 105)]
SPECAUG_KW = dict(freq_mask_num=2, time_mask_num=2)          # the fifth case uses two masks per axis


def specaug_wave(samples, seed):
    g = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(1, samples, generator=g)
