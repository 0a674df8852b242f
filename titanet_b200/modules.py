"""Drop-in for the reference's ``src/modules.py`` (same class names, constructor
signatures and ``state_dict`` keys), executing on libtitanet_sm100 kernels.

Public ``forward`` methods keep the reference's tensor contract (``[B, C, W]`` in and
out).  Inside the encoder the blocks talk to each other through :class:`Lazy`
(channels-last pre-BatchNorm tensors with the normalisation folded into the consumer's
load), which is what ``models.MegaBlock`` / ``models.Encoder`` use.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from . import _ops as ops

Tensor = torch.Tensor


@dataclass
class Lazy:
    """An activation in flight: ``a = dropout(relu(z * scale + shift))`` with
    ``z`` ``[B*T, C]`` channels-last.  ``scale is None`` => ``z`` already is the
    activation."""

    z: Tensor
    B: int
    T: int
    scale: Optional[Tensor] = None
    shift: Optional[Tensor] = None
    relu: bool = False
    p: float = 0.0
    seed: Optional[Tensor] = None
    layer: int = 0
    z2: Optional[Tensor] = None      # second alias of a plain ``z`` for its second consumer (ops.SETail2), or None

    @property
    def C(self) -> int:
        return self.z.shape[1]

    def materialise(self) -> "Lazy":
        if self.scale is None:
            return self
        if torch.is_grad_enabled():
            # two aliases (ops.Act2): a consumer pair (a mega-block's skip conv and first depthwise conv) sends its gradients
            # separately and tn_act_bwd2 adds them on load
            y, y2 = ops.Act2.apply(self.z, self.scale, self.shift, self.seed, self.relu, self.p, self.layer)
            return Lazy(y, self.B, self.T, z2=y2)
        y = ops.Act.apply(self.z, self.scale, self.shift, self.seed, self.relu, self.p, self.layer)
        return Lazy(y, self.B, self.T)

    def to_ncw(self) -> Tensor:
        m = self.materialise()
        return ops.nwc_to_ncw(m.z.view(m.B, m.T, m.C))

    @staticmethod
    def from_ncw(x: Tensor) -> "Lazy":
        if x.dim() != 3:
            raise ValueError(f"expected a [B, C, W] tensor, got {tuple(x.shape)}")
        B, C, T = x.shape
        return Lazy(ops.ncw_to_nwc(x).view(B * T, C), B, T)


class DropoutCtx:
    """Per-forward dropout bookkeeping: one device seed per forward pass (advanced by a
    kernel, so CUDA-graph replays draw fresh masks) and a distinct layer id per site."""

    def __init__(self, seed: Optional[Tensor]):
        self.seed = seed
        self._layer = 0

    def next_layer(self) -> int:
        self._layer += 1
        return self._layer


_GLOBAL_SEED_STATE = {}


def _seed_key(device) -> str:
    """One seed stream per physical device: 'cuda' and 'cuda:0' (the current device) are the same stream."""
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return str(dev)


def new_dropout_ctx(device, needed: bool) -> DropoutCtx:
    if not needed:
        return DropoutCtx(None)
    key = _seed_key(device)
    st = _GLOBAL_SEED_STATE.get(key)
    if st is None:
        st = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64).to(device)
        _GLOBAL_SEED_STATE[key] = st
    return DropoutCtx(ops.seed_next(st))


def reseed_dropout(seed: int, device="cuda"):
    """Reset the dropout seed stream of ``device`` (for reproducible runs): the next forward passes draw the same masks as
    the forward passes that followed the previous ``reseed_dropout(seed)``.  The state tensor is updated IN PLACE, so a
    captured CUDA graph (which holds its address) follows."""
    key = _seed_key(device)
    value = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64)
    st = _GLOBAL_SEED_STATE.get(key)
    if st is None:
        _GLOBAL_SEED_STATE[key] = value.to(torch.device(key))
    else:
        st.copy_(value)


def _check_conv_supported(conv: nn.Conv1d):
    if conv.stride[0] != 1 or conv.dilation[0] != 1:
        raise NotImplementedError("titanet_b200 convolutions support stride=1, dilation=1 only (all TitaNet uses)")
    if conv.kernel_size[0] % 2 == 0:
        raise NotImplementedError("titanet_b200 'same' convolutions need an odd kernel size")
    if conv.padding_mode != "zeros":
        raise NotImplementedError("only zero padding is supported")


class Conv1dSamePadding(nn.Conv1d):
    """1D convolution with "same" zero padding (reference: src/modules.py:5-40).
    Dense (groups=1) convs run as a conv-GEMM, groups == channels as the depthwise kernel."""

    def __init__(self, *args, **kwargs):
        super(Conv1dSamePadding, self).__init__(*args, **kwargs)

    def _fwd(self, x: Lazy, want_stats: bool = False):
        """Lazy in -> (z [B*T, Co], stats or None)."""
        _check_conv_supported(self)
        if self.groups == 1:
            x = x.materialise()
            return ops.conv_gemm(x.z, self.weight, self.bias, x.B, x.T, want_stats=want_stats)
        if self.groups == self.in_channels == self.out_channels:
            u = ops.Depthwise.apply(x.z, x.scale, x.shift, self.weight, self.bias, x.seed, x.relu, x.p, x.layer, x.B, x.T)
            stats = ops.ColStats.apply(u) if want_stats else None
            return u, stats
        raise NotImplementedError("grouped convolutions other than depthwise are not supported")

    def _fwd_bn(self, x: Lazy, bn: nn.BatchNorm1d):
        """Lazy in -> (z, scale, shift) of ``bn(conv(x))`` (BatchNorm folded, see ops.ConvGemmBN)."""
        _check_conv_supported(self)
        if self.groups == 1:
            x = x.materialise()
            if ops.conv_ktap_as_gemm_ok(x.z, self.weight, x.B, x.T):
                # K taps unrolled into the reduction dimension: the conv runs as one tensor-core GEMM (+ BatchNorm fold)
                K = self.weight.shape[2]
                kpad = (self.weight.shape[1] * K + 31) // 32 * 32
                x3 = ops.Im2Col.apply(x.z, x.B, x.T, K, kpad)
                w3 = ops.ConvWeightAsGemm.apply(self.weight, kpad)
                return ops.conv_gemm_bn(x3, w3, self.bias, bn, x.B, x.T)
            return ops.conv_gemm_bn(x.z, self.weight, self.bias, bn, x.B, x.T)
        z, stats = self._fwd(x, want_stats=bn.training)
        scale, shift = ops.bn_fold(stats, bn, float(z.shape[0]))
        return z, scale, shift

    def forward(self, inputs):
        x = Lazy.from_ncw(inputs)
        z, _ = self._fwd(x)
        return Lazy(z, x.B, x.T).to_ncw()


class DepthwiseConv1d(nn.Module):
    """Depthwise-separable convolution: depthwise K-tap then pointwise 1x1
    (reference: src/modules.py:43-93)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dilation=1, bias=True, device=None, dtype=None):
        super(DepthwiseConv1d, self).__init__()
        self.conv = nn.Sequential(
            Conv1dSamePadding(in_channels, in_channels, kernel_size=kernel_size, stride=stride, dilation=dilation,
                              groups=in_channels, bias=bias, device=device, dtype=dtype),
            Conv1dSamePadding(in_channels, out_channels, kernel_size=1, device=device, dtype=dtype),
        )

    def _fwd(self, x: Lazy, want_stats: bool = False):
        dw, pw = self.conv[0], self.conv[1]
        _check_conv_supported(dw)
        return ops.DwPw.apply(x.z, x.scale, x.shift, dw.weight, dw.bias, pw.weight, pw.bias, x.seed, x.relu, x.p, x.layer,
                              x.B, x.T, want_stats)

    def _fwd_bn(self, x: Lazy, bn: nn.BatchNorm1d):
        dw, pw = self.conv[0], self.conv[1]
        _check_conv_supported(dw)
        if ops._bn_trainable(bn):
            return ops.DwPwBN.apply(x.z, x.scale, x.shift, dw.weight, dw.bias, pw.weight, pw.bias, bn.weight, bn.bias,
                                    bn.running_mean, bn.running_var, bn.num_batches_tracked, bn.momentum, bn.eps, x.seed,
                                    x.relu, x.p, x.layer, x.B, x.T)
        z, stats = self._fwd(x, want_stats=bn.training)
        scale, shift = ops.bn_fold(stats, bn, float(z.shape[0]))
        return z, scale, shift

    def forward(self, inputs):
        x = Lazy.from_ncw(inputs)
        z, _ = self._fwd(x)
        return Lazy(z, x.B, x.T).to_ncw()


class ConvBlock1d(nn.Module):
    """conv -> BatchNorm1d -> activation -> dropout (reference: src/modules.py:96-148)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dilation=1, activation="relu", dropout=0,
                 depthwise=False):
        super(ConvBlock1d, self).__init__()
        assert activation is None or activation in ("relu", "tanh"), "Incompatible activation function"
        conv_module = DepthwiseConv1d if depthwise else Conv1dSamePadding
        modules = [
            conv_module(in_channels, out_channels, kernel_size=kernel_size, stride=stride, dilation=dilation),
            nn.BatchNorm1d(out_channels),
        ]
        if activation is not None:
            modules += [nn.ReLU() if activation == "relu" else nn.Tanh()]
        if dropout > 0:
            modules += [nn.Dropout(p=dropout)]
        self.conv_block = nn.Sequential(*modules)
        self._activation = activation
        self._dropout = float(dropout)

    def _fwd(self, x: Lazy, dctx: DropoutCtx) -> Lazy:
        """Lazy in -> Lazy out (pre-BN z plus the folded BN; ReLU/dropout deferred to the consumer)."""
        if self._activation == "tanh":
            raise NotImplementedError("ConvBlock1d(activation='tanh') is not on the TitaNet path and has no kernel")
        conv, bn = self.conv_block[0], self.conv_block[1]
        z, scale, shift = conv._fwd_bn(x, bn)
        p = self._dropout if self.training else 0.0
        return Lazy(z, x.B, x.T, scale, shift, relu=self._activation == "relu", p=p, seed=dctx.seed if p > 0 else None,
                    layer=dctx.next_layer())

    def forward(self, inputs):
        dctx = new_dropout_ctx(inputs.device, self.training and self._dropout > 0)
        return self._fwd(Lazy.from_ncw(inputs), dctx).to_ncw()


class SqueezeExcitation(nn.Module):
    """Squeeze-and-excitation gate (reference: src/modules.py:151-189).  Inside a
    ``MegaBlock`` it is fused with the residual tail (``models.MegaBlock``); standalone it
    is the same kernels with an all-zero skip branch."""

    def __init__(self, channels, reduction=16):
        super(SqueezeExcitation, self).__init__()
        self.squeeze = nn.AdaptiveAvgPool1d(1)
        self.excitation = nn.Sequential(
            nn.Linear(channels, channels // reduction, bias=False),
            nn.ReLU(),
            nn.Linear(channels // reduction, channels, bias=False),
            nn.Sigmoid(),
        )

    def forward(self, inputs):
        """[B, C, W] -> [B, C, W] (reference: src/modules.py:173-189): mean over time, two-layer gate, multiply."""
        lin1, lin2 = self.excitation[0], self.excitation[2]
        if lin1.bias is not None or lin2.bias is not None:
            raise NotImplementedError("SqueezeExcitation kernels have no bias terms (the reference uses bias=False)")
        x = Lazy.from_ncw(inputs)
        m = ops.MeanT.apply(x.z, x.B, x.T)
        gate = ops.SEMlp.apply(m, lin1.weight, lin2.weight)
        return Lazy(ops.GateMul.apply(x.z, gate, x.B, x.T), x.B, x.T).to_ncw()


class Squeeze(nn.Module):
    """Remove dimensions of size 1 (reference: src/modules.py:192-202); a view, no kernel."""

    def __init__(self, dim=None):
        super(Squeeze, self).__init__()
        self.dim = dim

    def forward(self, inputs):
        return inputs.squeeze(self.dim)
