"""Drop-in for the TitaNet part of the reference's ``src/models.py``: ``TitaNet``,
``Encoder``, ``MegaBlock``, ``Decoder``, ``AttentiveStatsPooling`` with the same
constructor signatures, class methods and ``state_dict`` keys, running on
libtitanet_sm100 kernels.  (``DumbConvNet`` and ``DVectorBaseline`` are outside the
accelerated path, see DESIGN.md.)
"""
from __future__ import annotations

from functools import partial

import numpy as np
import torch
import torch.nn as nn

from . import _ops as ops
from . import losses, modules
from ._lib import require_cuda
from .modules import DropoutCtx, Lazy, new_dropout_ctx


class TitaNet(nn.Module):
    """TitaNet speaker-embedding network (reference: src/models.py:162-339)."""

    TARGET_PARAMS = {"s": 6.4, "m": 13.4, "l": 25.3}

    def __init__(self, n_mels, n_mega_blocks, n_sub_blocks, encoder_hidden_size, encoder_output_size, embedding_size,
                 mega_block_kernel_size, prolog_kernel_size=3, epilog_kernel_size=1, attention_hidden_size=128,
                 se_reduction=16, simple_pool=False, loss_function=None, dropout=0.5, device="cpu"):
        super(TitaNet, self).__init__()
        self.encoder = Encoder(n_mels, n_mega_blocks, n_sub_blocks, encoder_hidden_size, encoder_output_size,
                               mega_block_kernel_size, prolog_kernel_size=prolog_kernel_size,
                               epilog_kernel_size=epilog_kernel_size, se_reduction=se_reduction, dropout=dropout)
        self.decoder = Decoder(encoder_output_size, attention_hidden_size, embedding_size, simple_pool=simple_pool)
        self.loss_function = loss_function
        self._dropout = float(dropout)
        self.to(device)

    def _tc_weights(self):
        """Weights of the 1x1 convs / linears that run on the tensor cores (pointwise, skip, epilog, ASP)."""
        ws = []
        for m in self.modules():
            if isinstance(m, nn.Conv1d) and m.groups == 1 and m.kernel_size[0] == 1:
                ws.append(m.weight)
            elif isinstance(m, AttentiveStatsPooling):
                ws += [m.in_linear.weight, m.out_linear.weight]
        return [w for w in ws if w.is_cuda and w.shape[0] % 32 == 0 and w.shape[1] % 32 == 0
                and (w.shape[0] % 128 == 0 or w.shape[1] % 128 == 0)]

    def _refresh_weight_splits(self):
        """One launch per step for every tf32 weight split (ops.SplitCache)."""
        if not ops.TC_ENABLED:
            return
        cache = self.__dict__.get("_split_cache")
        if cache is None or (cache and cache.stale()):
            weights = self._tc_weights()
            cache = ops.SplitCache(weights) if weights else False
            self.__dict__["_split_cache"] = cache
        if cache:
            cache.refresh()

    def get_n_params(self, div=1):
        """Number of trainable parameters, optionally divided (reference: src/models.py:221-228)."""
        return sum([np.prod(p.size()) for p in self.parameters() if p.requires_grad]) / div

    @classmethod
    def find_n_mega_blocks(cls, embedding_size, n_mels, model_size, loss_function=None, n_mega_blocks_trials=None):
        """Number of mega blocks whose parameter count is closest to the paper's
        (reference: src/models.py:230-260)."""
        if n_mega_blocks_trials is None:
            n_mega_blocks_trials = list(range(1, 20))
        target_params = cls.TARGET_PARAMS[model_size]
        best_value, min_distance = None, np.inf
        for n_mega_blocks in n_mega_blocks_trials:
            titanet = cls.get_titanet(embedding_size=embedding_size, n_mels=n_mels, n_mega_blocks=n_mega_blocks,
                                      model_size=model_size, loss_function=loss_function)
            params = titanet.get_n_params(div=1e6)
            distance = abs(target_params - params)
            if distance < min_distance:
                best_value = n_mega_blocks
                min_distance = distance
        return best_value

    @classmethod
    def get_titanet(cls, embedding_size=192, n_mels=80, n_mega_blocks=None, model_size="s", attention_hidden_size=128,
                    simple_pool=False, loss_function=None, dropout=0.5, device="cpu"):
        """TitaNet-S / -M / -L (reference: src/models.py:262-316)."""
        assert isinstance(model_size, str) and model_size.lower() in ("s", "m", "l"), "Unsupported model size"
        assert isinstance(loss_function, losses.MetricLearningLoss) or loss_function is None, "Unsupported loss function"
        if n_mega_blocks is None:
            n_mega_blocks = cls.find_n_mega_blocks(embedding_size, n_mels, model_size, loss_function=loss_function)
        titanet = partial(TitaNet, n_mels=n_mels, n_mega_blocks=n_mega_blocks, n_sub_blocks=3, encoder_output_size=1536,
                          embedding_size=embedding_size, attention_hidden_size=attention_hidden_size,
                          simple_pool=simple_pool, loss_function=loss_function, dropout=dropout, device=device)
        if model_size.lower() == "s":
            return titanet(encoder_hidden_size=256, mega_block_kernel_size=3)
        elif model_size.lower() == "m":
            return titanet(encoder_hidden_size=512, mega_block_kernel_size=7)
        elif model_size.lower() == "l":
            return titanet(encoder_hidden_size=1024, mega_block_kernel_size=11)

    def forward(self, spectrograms, speakers=None):
        """[B, M, T] spectrograms -> unit-norm embeddings [B, E]; with ``speakers`` the
        loss head's ``(embeddings, preds, loss)`` (reference: src/models.py:318-339)."""
        require_cuda(spectrograms)
        self._refresh_weight_splits()
        dctx = new_dropout_ctx(spectrograms.device, self.training and self._dropout > 0)
        encodings = self.encoder._fwd(Lazy.from_ncw(spectrograms), dctx)
        embeddings = self.decoder._fwd(encodings)
        if speakers is None:
            return ops.L2Norm.apply(embeddings, 1e-12)[0]
        assert self.loss_function is not None, "Loss function should not be None in training mode"
        return self.loss_function(embeddings, speakers)


class Encoder(nn.Module):
    """prolog -> MegaBlock x N -> epilog (reference: src/models.py:342-404)."""

    def __init__(self, n_mels, n_mega_blocks, n_sub_blocks, hidden_size, output_size, mega_block_kernel_size,
                 prolog_kernel_size=3, epilog_kernel_size=1, se_reduction=16, dropout=0.5):
        super(Encoder, self).__init__()
        self.prolog = modules.ConvBlock1d(n_mels, hidden_size, prolog_kernel_size)
        self.mega_blocks = nn.Sequential(
            *[MegaBlock(hidden_size, hidden_size, mega_block_kernel_size, n_sub_blocks, se_reduction=se_reduction,
                        dropout=dropout) for _ in range(n_mega_blocks)])
        self.epilog = modules.ConvBlock1d(hidden_size, output_size, epilog_kernel_size)
        self._dropout = float(dropout)

    def _fwd(self, x: Lazy, dctx: DropoutCtx) -> Lazy:
        x = self.prolog._fwd(x, dctx)
        for block in self.mega_blocks:
            x = block._fwd(x, dctx)
        return self.epilog._fwd(x, dctx)

    def forward(self, spectrograms):
        require_cuda(spectrograms)
        dctx = new_dropout_ctx(spectrograms.device, self.training and self._dropout > 0)
        return self._fwd(Lazy.from_ncw(spectrograms), dctx).to_ncw()


class MegaBlock(nn.Module):
    """sub-blocks (depthwise-separable conv, BN, ReLU, dropout) -> SE, plus a 1x1-conv/BN
    skip branch, ReLU and dropout (reference: src/models.py:407-472).

    Data flow per forward: one kernel group per sub-block writes the pre-BN tensor once
    (its BatchNorm statistics are accumulated in the GEMM epilogue; the next consumer
    applies BN + ReLU + dropout on load), the skip GEMM reads the block input, and one
    fused tail applies BN(skip) + SE gate * act + ReLU + dropout."""

    def __init__(self, input_size, output_size, kernel_size, n_sub_blocks, se_reduction=16, dropout=0.5):
        super(MegaBlock, self).__init__()
        self.dropout = dropout
        channels = [input_size] + [output_size] * n_sub_blocks
        self.sub_blocks = nn.Sequential(
            *[modules.ConvBlock1d(in_channels, out_channels, kernel_size, activation="relu", dropout=dropout,
                                  depthwise=True) for in_channels, out_channels in zip(channels[:-1], channels[1:])],
            modules.SqueezeExcitation(output_size, reduction=se_reduction))
        self.skip_connection = nn.Sequential(nn.Conv1d(input_size, output_size, kernel_size=1),
                                             nn.BatchNorm1d(output_size))

    def _fwd(self, x: Lazy, dctx: DropoutCtx) -> Lazy:
        x = x.materialise()
        B, T = x.B, x.T
        skip_conv, skip_bn = self.skip_connection[0], self.skip_connection[1]
        n_sub = len(self.sub_blocks) - 1
        first = self.sub_blocks[0] if n_sub > 0 else None
        fused_entry = (ops.FUSE_BLOCK_ENTRY and first is not None and isinstance(first.conv_block[0], modules.DepthwiseConv1d)
                       and first._activation == "relu" and ops._bn_trainable(first.conv_block[1]) and ops._bn_trainable(skip_bn)
                       and skip_conv.bias is not None and first.conv_block[0].conv[0].bias is not None)
        if fused_entry:
            # first sub-block + skip branch as one autograd node (ops.BlockEntryBN): no gradient-sum kernel on the block input
            dw, pw, bn1 = first.conv_block[0].conv[0], first.conv_block[0].conv[1], first.conv_block[1]
            modules._check_conv_supported(dw)
            z1, sc1, sh1, s, scs, shs = ops.BlockEntryBN.apply(
                x.z, dw.weight, dw.bias, pw.weight, pw.bias, bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var,
                bn1.num_batches_tracked, bn1.momentum, bn1.eps, skip_conv.weight, skip_conv.bias, skip_bn.weight, skip_bn.bias,
                skip_bn.running_mean, skip_bn.running_var, skip_bn.num_batches_tracked, skip_bn.momentum, skip_bn.eps, B, T)
            p1 = first._dropout if self.training else 0.0
            y = Lazy(z1, B, T, sc1, sh1, relu=True, p=p1, seed=dctx.seed if p1 > 0 else None, layer=dctx.next_layer())
        else:
            s, scs, shs = ops.conv_gemm_bn(x.z2 if x.z2 is not None else x.z, skip_conv.weight, skip_conv.bias, skip_bn, B, T)
            y = x
        for j in range(1 if fused_entry else 0, n_sub):
            y = self.sub_blocks[j]._fwd(y, dctx)
        if n_sub == 0 or y.scale is None:
            raise NotImplementedError("MegaBlock needs at least one sub-block")
        se = self.sub_blocks[n_sub]
        p_o = float(self.dropout) if self.training else 0.0
        seed = dctx.seed if (p_o > 0 or y.p > 0) else None
        args = (y.z, y.scale, y.shift, s, scs, shs, se.excitation[0].weight, se.excitation[2].weight, seed,
                y.p, y.layer, p_o, dctx.next_layer(), B, T)
        if ops.SE_TAIL_TWO_OUT and torch.is_grad_enabled():
            # two aliases of the block output: a following mega-block feeds its skip conv from the second one, so that the two
            # gradients reach SETail2.backward separately (summed on load) instead of through an autograd sum kernel
            out, out2 = ops.SETail2.apply(*args)
            return Lazy(out, B, T, z2=out2)
        return Lazy(ops.SETail.apply(*args), B, T)

    def forward(self, prolog_outputs):
        require_cuda(prolog_outputs)
        dctx = new_dropout_ctx(prolog_outputs.device, self.training and self.dropout > 0)
        return self._fwd(Lazy.from_ncw(prolog_outputs), dctx).to_ncw()


class Decoder(nn.Module):
    """Attentive statistics pooling -> BN -> Linear -> BN (reference: src/models.py:475-529)."""

    def __init__(self, encoder_output_size, attention_hidden_size, embedding_size, simple_pool=False):
        super(Decoder, self).__init__()
        if simple_pool:
            self.pool = nn.Sequential(nn.AdaptiveAvgPool1d(1), modules.Squeeze(-1),
                                      nn.Linear(encoder_output_size, encoder_output_size * 2))
        else:
            self.pool = nn.Sequential(AttentiveStatsPooling(encoder_output_size, attention_hidden_size),
                                      nn.BatchNorm1d(encoder_output_size * 2))
        self.linear = nn.Sequential(nn.Linear(encoder_output_size * 2, embedding_size), nn.BatchNorm1d(embedding_size))
        self._simple_pool = simple_pool

    def _fwd(self, enc: Lazy) -> torch.Tensor:
        B = enc.B
        if self._simple_pool:
            # AdaptiveAvgPool1d(1) -> Squeeze -> Linear(D, 2D)   (reference: src/models.py:497-502)
            x = enc.materialise()
            lin0 = self.pool[2]
            pooled, _ = ops.conv_gemm(ops.MeanT.apply(x.z, x.B, x.T), lin0.weight, lin0.bias, B, 1)
        else:
            pooled = self.pool[0]._fwd(enc)
            bn1 = self.pool[1]
            st1 = ops.ColStats.apply(pooled) if bn1.training else None
            sc1, sh1 = ops.bn_fold(st1, bn1, float(B))
            pooled = ops.Act.apply(pooled, sc1, sh1, None, False, 0.0, 0)
        lin, bn2 = self.linear[0], self.linear[1]
        z, sc2, sh2 = ops.conv_gemm_bn(pooled, lin.weight, lin.bias, bn2, B, 1)
        return ops.Act.apply(z, sc2, sh2, None, False, 0.0, 0)

    def forward(self, encodings):
        require_cuda(encodings)
        return self._fwd(Lazy.from_ncw(encodings))


class AttentiveStatsPooling(nn.Module):
    """Attention-weighted mean and standard deviation over time
    (reference: src/models.py:532-584)."""

    def __init__(self, input_size, hidden_size, eps=1e-6):
        super(AttentiveStatsPooling, self).__init__()
        self.eps = eps
        self.in_linear = nn.Linear(input_size, hidden_size)
        self.out_linear = nn.Linear(hidden_size, input_size)

    def _fwd(self, enc: Lazy) -> torch.Tensor:
        if enc.scale is not None:
            # the two consumers get their own alias of the materialised activation: no gradient-sum kernel (ops.Act2)
            xa, xb = ops.Act2.apply(enc.z, enc.scale, enc.shift, enc.seed, enc.relu, enc.p, enc.layer)
        else:
            xa = xb = enc.z
        h, _ = ops.conv_gemm(xa, self.in_linear.weight, self.in_linear.bias, enc.B, enc.T, tanh=True)
        e, _ = ops.conv_gemm(h, self.out_linear.weight, self.out_linear.bias, enc.B, enc.T)
        return ops.ASPPool.apply(e, xb, enc.B, enc.T, float(self.eps))

    def forward(self, encodings):
        require_cuda(encodings)
        return self._fwd(Lazy.from_ncw(encodings))
