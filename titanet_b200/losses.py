"""Drop-in for the reference's ``src/losses.py``: same class names, constructor signatures,
``state_dict`` keys (``fc.weight`` [, ``fc.bias``]) and return triple
``(normalised embeddings, predictions, loss)``; the arithmetic runs on libtitanet_sm100
kernels (linear layer as a conv-GEMM, one fused row kernel per loss for forward and
gradient)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _ops as ops


class MetricLearningLoss(nn.Module):
    """Generic loss function to be used in a metric learning setting
    (reference: src/losses.py:7-19)."""

    def __init__(self, embedding_size, n_classes, device="cpu", *args, **kwargs):
        super(MetricLearningLoss, self).__init__()
        self.embedding_size = embedding_size
        self.n_classes = n_classes
        self.device = device

    def forward(self, inputs, targets):
        raise NotImplementedError()


class CELoss(MetricLearningLoss):
    """Linear layer + cross-entropy (reference: src/losses.py:22-44)."""

    def __init__(self, embedding_size, n_classes, device="cpu"):
        super(CELoss, self).__init__(embedding_size, n_classes, device=device)
        self.fc = nn.Linear(embedding_size, n_classes)

    def forward(self, inputs, targets):
        B = inputs.shape[0]
        logits, _ = ops.conv_gemm(inputs, self.fc.weight, self.fc.bias, B, 1)
        loss, preds = ops.CrossEntropy.apply(logits, targets)
        normalised, _ = ops.L2Norm.apply(inputs, 1e-12)
        return normalised, preds, loss


class AngularMarginLoss(MetricLearningLoss):
    """Generic angular margin loss (reference: src/losses.py:47-132).

    Kept quirks: the class weights are row-normalised *in place* on every forward,
    outside autograd (line 86), so the gradient is taken w.r.t. the normalised weights;
    inputs are divided by their norm without eps (89-92); ``scale=None`` uses the
    per-sample input norm (95-99); no max-subtraction before ``exp`` (127); ``eps`` is
    added to the denominator (130)."""

    def __init__(self, embedding_size, n_classes, device="cpu", scale=None, m1=1, m2=0, m3=0, eps=1e-6):
        super(AngularMarginLoss, self).__init__(embedding_size, n_classes, device=device)
        self.fc = nn.Linear(embedding_size, n_classes, bias=False)
        self.scale = scale
        self.m1 = m1
        self.m2 = m2
        self.m3 = m3
        self.eps = eps

    def forward(self, inputs, targets):
        B = inputs.shape[0]
        ops.rownorm_(self.fc.weight.data)
        normalised, norms = ops.L2Norm.apply(inputs, 0.0)
        raw, _ = ops.conv_gemm(normalised, self.fc.weight, None, B, 1)
        loss, preds = ops.AngularMargin.apply(raw, norms if self.scale is None else None, targets, self.scale,
                                              float(self.m1), float(self.m2), float(self.m3), float(self.eps))
        return normalised, preds, loss


class SphereFaceLoss(AngularMarginLoss):
    """Multiplicative angular margin (reference: src/losses.py:135-149)."""

    def __init__(self, embedding_size, n_classes, device="cpu", scale=None, margin=3, eps=1e-6):
        assert margin > 1, "Margin out of bounds"
        super(SphereFaceLoss, self).__init__(embedding_size, n_classes, device=device, scale=scale, m1=margin, eps=eps)


class CosFaceLoss(AngularMarginLoss):
    """Additive cosine margin (reference: src/losses.py:152-166)."""

    def __init__(self, embedding_size, n_classes, device="cpu", scale=64, margin=0.2, eps=1e-6):
        assert margin > 0 and margin < 1 - math.cos(math.pi / 4), "Margin out of bounds"
        super(CosFaceLoss, self).__init__(embedding_size, n_classes, device=device, scale=scale, m3=margin, eps=eps)


class ArcFaceLoss(AngularMarginLoss):
    """Additive angular margin (reference: src/losses.py:169-183)."""

    def __init__(self, embedding_size, n_classes, device="cpu", scale=64, margin=0.5, eps=1e-6):
        assert margin > 0 and margin < 1, "Margin out of bounds"
        super(ArcFaceLoss, self).__init__(embedding_size, n_classes, device=device, scale=scale, m2=margin, eps=eps)


class GE2ELoss(MetricLearningLoss):
    """GE2E (reference: src/losses.py:186-261) is outside the accelerated path (a Python
    triple loop over speakers in the reference, not named by any benchmark config); the
    class exists so ``LOSSES['ge2e']`` resolves, and says so when called."""

    def __init__(self, embedding_size, n_classes, device="cpu"):
        super(GE2ELoss, self).__init__(embedding_size, n_classes, device=device)
        self.w = nn.Parameter(torch.tensor(1.0))
        self.b = nn.Parameter(torch.tensor(0.0))

    def forward(self, inputs, targets):
        raise NotImplementedError("GE2ELoss has no titanet_b200 kernel (out of the hot-path scope, see DESIGN.md)")


LOSSES = {
    "ce": CELoss,
    "sphere": SphereFaceLoss,
    "cos": CosFaceLoss,
    "arc": ArcFaceLoss,
    "ge2e": GE2ELoss,
}
