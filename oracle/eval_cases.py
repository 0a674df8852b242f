"""Seeded trial lists / utterance sets shared by ``oracle/make_golden_eval.py`` and ``tests/``.  TEST INFRASTRUCTURE."""
import numpy as np
import torch

from cases import TINY


def trial_cases():
    """name -> (scores fp32 [n], labels int64 [n]).  Target scores sit higher than non-target ones, with overlap."""
    out = {}
    rng = np.random.RandomState(1234)

    def draw(n, frac, sep, decimals=None):
        lab = (rng.rand(n) < frac).astype(np.int64)
        s = (rng.randn(n) * 0.25 + sep * lab).astype(np.float32)
        s = np.clip(s, -1, 1)
        if decimals is not None:
            s = np.round(s, decimals).astype(np.float32)
        return s, lab

    out["random_200"] = draw(200, 0.3, 0.4)
    out["ties_500"] = draw(500, 0.2, 0.3, decimals=1)             # ~20 distinct scores: ROC has diagonal segments
    out["ties_3000"] = draw(3000, 0.1, 0.5, decimals=2)           # > one 2048-chunk of the device sort
    out["tiny_5"] = (np.asarray([0.9, 0.1, 0.4, 0.35, 0.8], np.float32), np.asarray([1, 0, 1, 0, 0], np.int64))
    out["separable_64"] = (np.r_[np.linspace(0.6, 0.9, 20), np.linspace(-0.5, 0.3, 44)].astype(np.float32),
                           np.r_[np.ones(20), np.zeros(44)].astype(np.int64))
    out["signed_zero_40"] = (np.asarray([0.0, -0.0, 0.25, -0.25] * 10, np.float32),
                             np.asarray([1, 0, 0, 1, 1, 0, 1, 0] * 5, np.int64))
    out["large_20000"] = draw(20000, 0.05, 0.45)                  # two global bitonic stages
    return out


MODEL_SPEC = TINY["tiny_k3"]
MODEL_FRAMES = [50, 37, 50, 64, 37, 50, 41]
MODEL_SPEAKERS = [3, 7, 3, 9, 7, 9, 3]


def model_utterances(seed=11):
    """A tiny test split: 7 spectrograms ``[1, n_mels, T_i]`` of 4 different lengths, 3 speakers."""
    g = torch.Generator().manual_seed(seed)
    return [0.3 * torch.randn(1, MODEL_SPEC.n_mels, t, generator=g) for t in MODEL_FRAMES]
