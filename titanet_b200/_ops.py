"""Autograd bindings of the libtitanet_sm100 kernels.

Internal tensor convention: activations are channels-last ``[B*T, C]`` fp32 ("NWC").
A tensor in front of a BatchNorm travels as its pre-BN values ``z`` plus the folded
per-channel ``(scale, shift)`` -- the consumer kernel applies affine + ReLU + dropout
while loading (see include/titanet_b200.h, "lazy activation").  BatchNorm statistics
are an ordinary differentiable fp64 tensor ``stats = [sum | sum of squares]`` produced by
the conv-GEMM epilogue; its gradient is folded back into ``dz`` by ``tn_stats_bwd``.

Every op here launches hand-written CUDA kernels through the C ABI; nothing falls back
to torch math.  torch is used for memory (``torch.empty``), streams and autograd only.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch.autograd import Function

import ctypes
import weakref

from ._lib import LIB, TnBnBwd, TnBnFold, TnScratch, TnSplitJob, call, ptr, require_cuda, stream

Tensor = torch.Tensor
EPI_TANH, EPI_ACCUM, GEMM_GRAD = 1, 2, 8
TN_TICKETS = 64


# ----------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------
def _c(t: Optional[Tensor]) -> Optional[Tensor]:
    """contiguous fp32 CUDA tensor (or None)."""
    if t is None:
        return None
    require_cuda(t)
    if t.dtype != torch.float32:
        raise TypeError(f"titanet_b200 kernels are fp32; got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def empty(shape, ref: Tensor, dtype=torch.float32) -> Tensor:
    return torch.empty(shape, device=ref.device, dtype=dtype)


class ZeroArena:
    """One zero-filled device buffer per step for every accumulator the kernels add into (BatchNorm
    statistics, per-channel reductions, every parameter gradient): ``reset()`` is ONE memset, ``take``
    carves 256-byte aligned views.  A TitaNet-S step has ~500 such buffers; as separate memset nodes
    they cost more than the arithmetic they serve.  Used by ``engine.GraphedTrainStep`` (fixed shapes,
    so the carve-up is identical on every step); outside an arena ``zeros`` is a plain per-tensor memset.

    ``grads_only`` views handed out by ``gempty`` (fully overwritten gradients) live in the same buffer,
    so after backward every ``p.grad`` is a view of ``self.buf`` and the data-parallel exchange can
    all-reduce the arena in place."""

    ALIGN = 256

    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:      # tensors report an explicit index
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.buf: Optional[Tensor] = None
        self.offset = 0
        self.high_water = 0
        self.measuring = True

    def begin_step(self):
        """Start a step: zero the arena (allocated after the first, measuring, step)."""
        if self.buf is None and self.high_water > 0 and not self.measuring:
            self.buf = torch.empty(self.high_water, device=self.device, dtype=torch.uint8)
        self.offset = 0
        if self.buf is not None:
            LIB.call("tn_zero", self.buf.data_ptr(), self.buf.numel(), stream())

    def finish_measuring(self):
        self.measuring = False

    def take(self, shape, dtype) -> Optional[Tensor]:
        n = 1
        for d in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)):
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        start = self.offset
        self.offset = (start + nbytes + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if self.buf is None:
            self.high_water = max(self.high_water, self.offset)
            return None
        if self.offset > self.buf.numel():
            raise RuntimeError("ZeroArena overflow: the step's shapes changed after the arena was sized")
        return self.buf[start:start + nbytes].view(dtype).view(shape)


_ARENA: Optional[ZeroArena] = None


def set_arena(arena: Optional[ZeroArena]) -> Optional[ZeroArena]:
    """Install (or remove, with None) the arena ``zeros`` / ``gempty`` carve from; returns the previous one."""
    global _ARENA
    prev, _ARENA = _ARENA, arena
    return prev


def zeros(shape, ref: Tensor, dtype=torch.float32) -> Tensor:
    if _ARENA is not None and ref.device == _ARENA.device:
        t = _ARENA.take(shape, dtype)
        if t is not None:
            return t
    t = torch.empty(shape, device=ref.device, dtype=dtype)
    if t.numel():
        LIB.call("tn_zero", t.data_ptr(), t.numel() * t.element_size(), stream())
    return t


def gempty(shape, ref: Tensor, dtype=torch.float32) -> Tensor:
    """Uninitialised storage for a parameter gradient the kernel fully overwrites (inside the arena when
    one is active, so that all gradients share one buffer)."""
    if _ARENA is not None and ref.device == _ARENA.device:
        t = _ARENA.take(shape, dtype)
        if t is not None:
            return t
    return torch.empty(shape, device=ref.device, dtype=dtype)


def seed_next(state: Tensor) -> Tensor:
    """Advance the int64[1] device seed state and return this step's seed tensor."""
    require_cuda(state)
    out = torch.empty(1, device=state.device, dtype=torch.int64)
    call("tn_seed_next", ptr(state), ptr(out))
    return out


_TICKETS = {}
ACCUM_MAX_CHANNELS = 8192


def scratch(ref: Tensor, floats: int = 0):
    """(ctypes ``tn_scratch``, keep-alive tensor) for the atomics-free cross-block reductions of the conv-GEMM entry points:
    the zero-initialised ticket array and fixed-point statistics accumulators of the current stream (kernels return both to
    zero, so one pair serves every call on a stream, CUDA-graph replays included) plus, when ``floats`` > 0, an
    uninitialised fp32 workspace for split-K partial tiles."""
    key = (ref.device, stream())
    ent = _TICKETS.get(key)
    if ent is None:
        words = 4 * ACCUM_MAX_CHANNELS + TN_TICKETS // 2
        tk = torch.empty(TN_TICKETS, device=ref.device, dtype=torch.int32)
        acc = torch.empty(words, device=ref.device, dtype=torch.int64)
        LIB.call("tn_zero", tk.data_ptr(), tk.numel() * 4, stream())
        LIB.call("tn_zero", acc.data_ptr(), acc.numel() * 8, stream())
        ent = _TICKETS[key] = (tk, acc)
    tk, acc = ent
    parts = torch.empty(int(floats), device=ref.device, dtype=torch.float32) if floats > 0 else None
    return TnScratch(ptr(parts), int(floats), tk.data_ptr(), acc.data_ptr(), acc.numel()), parts


# ----------------------------------------------------------------------------
# layout
# ----------------------------------------------------------------------------
class Transpose(Function):
    """[B, C, T] -> [B, T, C] (to_nwc) or back."""

    @staticmethod
    def forward(ctx, x: Tensor, to_nwc: bool):
        x = _c(x)
        ctx.to_nwc = to_nwc
        if to_nwc:
            B, C, T = x.shape
            y = empty((B, T, C), x)
            call("tn_ncw_to_nwc", ptr(x), ptr(y), B, C, T)
        else:
            B, T, C = x.shape
            y = empty((B, C, T), x)
            call("tn_nwc_to_ncw", ptr(x), ptr(y), B, C, T)
        return y

    @staticmethod
    def backward(ctx, dy):
        return Transpose.apply(dy, not ctx.to_nwc), None


def ncw_to_nwc(x: Tensor) -> Tensor:
    return Transpose.apply(x, True)


def nwc_to_ncw(x: Tensor) -> Tensor:
    return Transpose.apply(x, False)


# ----------------------------------------------------------------------------
# conv / linear as GEMM  (+ BatchNorm statistics in the epilogue)
# ----------------------------------------------------------------------------
# Tensor-core (tcgen05) path for the 1x1 convs / linears.  TC_FWD_NSPLIT / TC_BWD_NSPLIT:
# 3 = fp32-equivalent split (TF32 + one bf16 correction MMA; 3xTF32 under TN_TC_3XTF32=1), 1 = plain TF32.  TC_ENABLED exists for A/B tests against the
# exact-fp32 CUDA-core kernel, not as a runtime fallback (unsupported shapes always take the
# CUDA-core kernel: K-tap convs, channel counts that are not multiples of 128 / 32).
TC_ENABLED = True
TC_FWD_NSPLIT = 3
TC_BWD_NSPLIT = 3
TC_WGRAD = True
TC_MIN_ROWS = 256
WS_PLANES = 4      # include/titanet_b200.h TN_WS_PLANES: tf32 hi | tf32 lo | bf16 correction rows | scaled-fp16 correction rows


def _tc_ok(R: int, Kd: int, M: int, K: int) -> bool:
    return TC_ENABLED and K == 1 and R >= TC_MIN_ROWS and Kd % 32 == 0 and M % 128 == 0


class SplitCache:
    """The tf32 (hi | lo) splits of every tensor-core weight of a model, both orientations, produced by
    ONE kernel per step (``tn_split_tf32_batch``) into a persistent buffer instead of one tiny launch in
    front of every GEMM (142 per TitaNet-S step).  ``refresh()`` is called at the top of the model's
    forward; ``lookup`` is valid for weights whose ``_version`` has not moved since."""

    def __init__(self, weights):
        self.weights = [w for w in weights]
        dev = self.weights[0].device
        self.ptrs = [w.data_ptr() for w in self.weights]
        jobs, self.views, off, tiles = [], {}, 0, 0
        total = sum(2 * WS_PLANES * w.numel() for w in self.weights)
        self.buf = torch.empty(total, device=dev, dtype=torch.float32)
        for i, w in enumerate(self.weights):
            Co, Ci = w.shape[0], w.shape[1]
            n = Co * Ci
            fwd = self.buf[off:off + WS_PLANES * n].view(WS_PLANES, Co, Ci)
            bwd = self.buf[off + WS_PLANES * n:off + 2 * WS_PLANES * n].view(WS_PLANES, Ci, Co)
            off += 2 * WS_PLANES * n
            # tile0: the job's first 32 x 32 tile in the step's tile list (tn_split_tf32_batch)
            jobs.append(TnSplitJob(w.data_ptr(), fwd.data_ptr(), Co, Ci, 0, tiles))
            tiles += -(-Co // 32) * (Ci // 32)
            jobs.append(TnSplitJob(w.data_ptr(), bwd.data_ptr(), Ci, Co, 1, tiles))
            tiles += -(-Ci // 32) * (Co // 32)
            self.views[id(w)] = (i, fwd, bwd)
        arr = (TnSplitJob * len(jobs))(*jobs)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.jobs = raw.to(dev)
        self.njobs = len(jobs)
        self.total_tiles = tiles
        self.versions = [-1] * len(self.weights)

    def stale(self) -> bool:
        return any(w.data_ptr() != p for w, p in zip(self.weights, self.ptrs))

    def refresh(self):
        call("tn_split_tf32_batch", ptr(self.jobs), self.njobs, self.total_tiles)
        self.versions = [w._version for w in self.weights]
        for w in self.weights:
            _SPLITS[id(w)] = self

    def lookup(self, w: Tensor):
        ent = self.views.get(id(w))
        if ent is None or self.weights[ent[0]] is not w or self.versions[ent[0]] != w._version \
                or self.ptrs[ent[0]] != w.data_ptr():
            return None
        return ent[1], ent[2]


# id(parameter) -> its SplitCache.  Weak values: the cache is owned by the model (which also keeps the parameters, and
# with them their ids, alive), so entries vanish with the model instead of pinning its buffers.
_SPLITS = weakref.WeakValueDictionary()


def cached_splits(w: Tensor):
    """(ws_fwd [WS_PLANES, Co, Ci], ws_dgrad [WS_PLANES, Ci, Co]) of a weight refreshed this step, else None."""
    cache = _SPLITS.get(id(w))
    return cache.lookup(w) if cache is not None else None


def make_bn_fold(gamma, beta, rm, rv, nbt, momentum, eps, n, scale, shift, mean, invstd) -> TnBnFold:
    return TnBnFold(ptr(gamma), ptr(beta), ptr(rm), ptr(rv), ptr(nbt), float(momentum), float(eps), float(n), ptr(scale),
                    ptr(shift), ptr(mean), ptr(invstd))


def _gemm_tc(x, w2, bias, z, stats, R, Kd, M, transpose, flags, nsplit, tag, ws=None, bn=None):
    if ws is None:
        ws = torch.empty((WS_PLANES, M, Kd), device=x.device, dtype=torch.float32)
        call("tn_split_tf32", ptr(w2), ptr(ws), M, Kd, int(transpose))
    sc, keep = scratch(x) if stats is not None else (None, None)
    scp = ctypes.byref(sc) if sc is not None else None
    if bn is not None:
        call("tn_gemm_tc_bn", ptr(x), ptr(ws), ptr(bias), ptr(z), ptr(stats), ctypes.byref(bn), R, Kd, M, flags, nsplit, scp, tag=tag)
    else:
        call("tn_gemm_tc", ptr(x), ptr(ws), ptr(bias), ptr(z), ptr(stats), R, Kd, M, flags, nsplit, scp, tag=tag)


def gemm_tc_raw(x, ws, bias, z, stats, R, Kd, M, flags, nsplit):
    """``tn_gemm_tc`` on raw tensors with the scratch it needs (tools / kernel tests)."""
    sc, keep = scratch(x) if stats is not None else (None, None)
    call("tn_gemm_tc", ptr(x), ptr(ws), ptr(bias), ptr(z), ptr(stats), R, Kd, M, flags, nsplit,
         ctypes.byref(sc) if sc is not None else None)


def _gemm_fwd(x, w3, bias, z, stats, B, T, transpose_w, flags, ws=None, bn=None):
    """ws: pre-split weights for this orientation (SplitCache) or None; bn: TnBnFold to fuse (forward only)."""
    Co, Ci, K = w3.shape
    R = B * T
    if transpose_w and _tc_ok(R, Co, Ci, K):
        return _gemm_tc(x, w3, bias, z, stats, R, Co, Ci, 1, flags | GEMM_GRAD, TC_BWD_NSPLIT, f"dgrad R{R} Ci{Co} Co{Ci} K1", ws=ws)
    if not transpose_w and _tc_ok(R, Ci, Co, K):
        return _gemm_tc(x, w3, bias, z, stats, R, Ci, Co, 0, flags, TC_FWD_NSPLIT, f"fwd R{R} Ci{Ci} Co{Co} K1", ws=ws, bn=bn)
    co_out, ci_red = (Ci, Co) if transpose_w else (Co, Ci)
    need = LIB.query("tn_conv_gemm_simt_scratch_floats", B, T, ci_red, co_out, K, flags)
    sc, keep = scratch(x, need) if (need > 0 or stats is not None) else (None, None)
    scp = ctypes.byref(sc) if sc is not None else None
    if transpose_w:
        call("tn_conv_gemm_simt", ptr(x), ptr(w3), ptr(bias), ptr(z), ptr(stats), B, T, Co, Ci, K, 1, flags, scp,
             tag=f"dgrad R{B * T} Ci{Co} Co{Ci} K{K}")
    elif bn is not None:
        call("tn_conv_gemm_simt_bn", ptr(x), ptr(w3), ptr(bias), ptr(z), ptr(stats), ctypes.byref(bn), B, T, Ci, Co, K, flags, scp,
             tag=f"fwd R{B * T} Ci{Ci} Co{Co} K{K}")
    else:
        call("tn_conv_gemm_simt", ptr(x), ptr(w3), ptr(bias), ptr(z), ptr(stats), B, T, Ci, Co, K, 0, flags, scp,
             tag=f"fwd R{B * T} Ci{Ci} Co{Co} K{K}")


def _gemm_wgrad(dz, x, dw3, dbias, B, T):
    Co, Ci, K = dw3.shape
    R = B * T
    if TC_ENABLED and TC_WGRAD and K == 1 and R >= TC_MIN_ROWS and Ci % 32 == 0 and Co % 128 == 0:
        call("tn_wgrad_tc", ptr(dz), ptr(x), ptr(dw3), R, Ci, Co, tag=f"wgrad R{R} Ci{Ci} Co{Co} K1")
        if dbias is not None:
            call("tn_colsum", ptr(dz), ptr(dbias), R, Co)
        return
    call("tn_conv_wgrad_simt", ptr(dz), ptr(x), ptr(dw3), ptr(dbias), B, T, Ci, Co, K, tag=f"wgrad R{B * T} Ci{Ci} Co{Co} K{K}")


class ConvGemm(Function):
    """z[B*T, Co] = conv1d_same(x[B*T, Ci], w[Co, Ci, K]) + bias, optional tanh epilogue,
    optional BatchNorm statistics of z.  ``w`` may be 2-D (a linear layer)."""

    @staticmethod
    def forward(ctx, x, w, bias, B: int, T: int, want_stats: bool, tanh: bool):
        x, w, bias = _c(x), _c(w), _c(bias)
        w3 = w if w.dim() == 3 else w.unsqueeze(-1)
        Co, Ci, K = w3.shape
        assert x.shape == (B * T, Ci), (x.shape, B, T, Ci)
        z = empty((B * T, Co), x)
        stats = empty((2 * Co,), x, torch.float64) if want_stats else None
        sp = cached_splits(w)
        _gemm_fwd(x, w3, bias, z, stats, B, T, 0, EPI_TANH if tanh else 0, ws=sp[0] if sp else None)
        ctx.ws_t = sp[1] if sp else None
        ctx.save_for_backward(x, w, bias, z if (want_stats or tanh) else None)
        ctx.meta = (B, T, want_stats, tanh)
        if want_stats:
            return z, stats
        return z, None

    @staticmethod
    def backward(ctx, dz, dstats):
        x, w, bias, z = ctx.saved_tensors
        B, T, want_stats, tanh = ctx.meta
        w3 = w if w.dim() == 3 else w.unsqueeze(-1)
        Co, Ci, K = w3.shape
        if dz is None:
            dz = zeros((B * T, Co), x)
        dz = _c(dz)
        if tanh:
            g = empty(dz.shape, dz)
            call("tn_tanh_bwd", ptr(dz), ptr(z), ptr(g), dz.numel())
            dz = g
        db = zeros(bias.shape, bias) if bias is not None else None
        db_done = False
        if want_stats and dstats is not None:
            g = empty(dz.shape, dz)
            call("tn_stats_bwd", ptr(dz), ptr(z), ptr(dstats.contiguous()), ptr(g), ptr(db), B * T, Co)   # + bias gradient
            dz, db_done = g, True
        dx = None
        if ctx.needs_input_grad[0]:
            dx = empty((B * T, Ci), x)
            _gemm_fwd(dz, w3, None, dx, None, B, T, 1, 0, ws=ctx.ws_t)
        dw = zeros(w.shape, w)
        _gemm_wgrad(dz, x, dw if dw.dim() == 3 else dw.unsqueeze(-1), None if db_done else db, B, T)
        return dx, dw, db, None, None, None, None


def conv_gemm(x, w, bias, B, T, want_stats=False, tanh=False):
    return ConvGemm.apply(x, w, bias, B, T, want_stats, tanh)


def _bn_forward_buffers(Co: int, ref: Tensor):
    """(stats fp64 [2Co] (written by the kernel), fold [4, Co] = scale | shift | mean | invstd)."""
    stats = empty((2 * Co,), ref, torch.float64)
    fold = empty((4, Co), ref)
    return stats, fold


def _bn_backward(dz, z, dscale, dshift, fold, gamma, n, bias, R, Co):
    """tn_bn_stats_bwd: returns (g = full gradient w.r.t. z, dbias, dgamma, dbeta)."""
    dscale = _c(dscale) if dscale is not None else zeros((Co,), z)
    dshift = _c(dshift) if dshift is not None else zeros((Co,), z)
    g = empty(z.shape, z)
    db = zeros((Co,), z) if bias is not None else None
    dgamma, dbeta = gempty((Co,), z), gempty((Co,), z)
    call("tn_bn_stats_bwd", ptr(_c(dz)) if dz is not None else None, ptr(z), ptr(dscale), ptr(dshift), fold[2].data_ptr(),
         fold[3].data_ptr(), ptr(gamma), float(n), ptr(g), ptr(db), ptr(dgamma), ptr(dbeta), R, Co)
    return g, db, dgamma, dbeta


# BatchNorm backward folded into the data-gradient GEMM's operand load (tn_gemm_tc_bnbwd / tn_gemm_tc_dwbwd_bn): the transform
# warps of the pair kernel build g = dz + a[c] + b[c] z on the fly and write it out once for the weight-gradient GEMM, so the
# tn_bn_stats_bwd launch (read dz, z; write g) disappears.  TN_FUSE_BNBWD=0 restores the separate kernel (A/B).
TC_FUSE_BNBWD = __import__("os").environ.get("TN_FUSE_BNBWD", "1") != "0"


def _bnbwd_fusable(dz, R: int, Kd: int, M: int) -> bool:
    return (TC_ENABLED and TC_FUSE_BNBWD and TC_BWD_NSPLIT == 3 and dz is not None and R >= 512 and M % 256 == 0 and Kd % 32 == 0
            and M % 128 == 0)


def _make_bn_bwd(dz, z, dscale, dshift, fold, gamma, n, bias, Co):
    """(TnBnBwd, g, dbias, dgamma, dbeta, keep-alive list) for a fused BatchNorm backward."""
    dscale = _c(dscale) if dscale is not None else zeros((Co,), z)
    dshift = _c(dshift) if dshift is not None else zeros((Co,), z)
    g = empty(z.shape, z)
    db = zeros((Co,), z) if bias is not None else None
    dgamma, dbeta = gempty((Co,), z), gempty((Co,), z)
    bnb = TnBnBwd(ptr(z), ptr(dscale), ptr(dshift), fold[2].data_ptr(), fold[3].data_ptr(), ptr(gamma), float(n), ptr(g), ptr(db),
                  ptr(dgamma), ptr(dbeta))
    return bnb, g, db, dgamma, dbeta, (dscale, dshift)


class ConvGemmBN(Function):
    """conv (as GEMM) followed by a TRAIN-mode BatchNorm1d, folded: returns the pre-BN tensor ``z`` and the
    per-channel ``(scale, shift)``.  The GEMM epilogue accumulates the statistics and its last CTA folds them
    (and updates running_mean / running_var / num_batches_tracked), so there is no separate BatchNorm launch;
    backward is one pass (tn_bn_stats_bwd) + dgrad + wgrad.  (src/modules.py:119-131, src/models.py:452-455)"""

    @staticmethod
    def forward(ctx, x, w, bias, gamma, beta, rm, rv, nbt, momentum: float, eps: float, B: int, T: int):
        x, w, bias, gamma, beta = _c(x), _c(w), _c(bias), _c(gamma), _c(beta)
        w3 = w if w.dim() == 3 else w.unsqueeze(-1)
        Co, Ci, K = w3.shape
        assert x.shape == (B * T, Ci), (x.shape, B, T, Ci)
        z = empty((B * T, Co), x)
        stats, fold = _bn_forward_buffers(Co, x)
        n = float(B * T)
        bn = make_bn_fold(gamma, beta, rm, rv, nbt, momentum, eps, n, fold[0], fold[1], fold[2], fold[3])
        sp = cached_splits(w)
        _gemm_fwd(x, w3, bias, z, stats, B, T, 0, 0, ws=sp[0] if sp else None, bn=bn)
        ctx.ws_t = sp[1] if sp else None
        ctx.save_for_backward(x, w, bias, z, gamma, fold)
        ctx.meta = (B, T, n)
        return z, fold[0], fold[1]

    @staticmethod
    def backward(ctx, dz, dscale, dshift):
        x, w, bias, z, gamma, fold = ctx.saved_tensors
        B, T, n = ctx.meta
        w3 = w if w.dim() == 3 else w.unsqueeze(-1)
        Co, Ci, K = w3.shape
        R = B * T
        if ctx.needs_input_grad[0] and K == 1 and _bnbwd_fusable(dz, R, Co, Ci):
            bnb, g, db, dgamma, dbeta, keep = _make_bn_bwd(_c(dz), z, dscale, dshift, fold, gamma, n, bias, Co)
            ws = ctx.ws_t
            if ws is None:
                ws = torch.empty((WS_PLANES, Ci, Co), device=x.device, dtype=torch.float32)
                call("tn_split_tf32", ptr(w3), ptr(ws), Ci, Co, 1)
            dx = empty((R, Ci), x)
            call("tn_gemm_tc_bnbwd", ptr(_c(dz)), ptr(ws), ctypes.byref(bnb), ptr(dx), R, Co, Ci, 0, tag=f"bnbwd+dgrad R{R} Ci{Co} Co{Ci} K1")
        else:
            g, db, dgamma, dbeta = _bn_backward(dz, z, dscale, dshift, fold, gamma, n, bias, R, Co)
            dx = None
            if ctx.needs_input_grad[0]:
                dx = empty((R, Ci), x)
                _gemm_fwd(g, w3, None, dx, None, B, T, 1, 0, ws=ctx.ws_t)
        dw = zeros(w.shape, w)
        _gemm_wgrad(g, x, dw if dw.dim() == 3 else dw.unsqueeze(-1), None, B, T)
        return dx, dw, db, dgamma, dbeta, None, None, None, None, None, None, None


class Im2Col(Function):
    """X3[r, tap * Ci + ci] = x[r + tap - K/2, ci] (zero outside the utterance / up to Kpad): a K-tap dense conv becomes the
    1x1 GEMM X3 W3^T on the tensor cores.  No input gradient (the caller takes the CUDA-core path when the input needs one)."""

    @staticmethod
    def forward(ctx, x, B: int, T: int, K: int, Kpad: int):
        x = _c(x)
        out = empty((B * T, Kpad), x)
        call("tn_im2col_nwc", ptr(x), ptr(out), B, T, x.shape[1], K, Kpad)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        return None, None, None, None, None


class ConvWeightAsGemm(Function):
    """w[Co, Ci, K] -> w3[Co, Kpad, 1] with w3[co, tap * Ci + ci] = w[co, ci, tap]; backward reorders the gradient back."""

    @staticmethod
    def forward(ctx, w, Kpad: int):
        w = _c(w)
        Co, Ci, K = w.shape
        w3 = empty((Co, Kpad, 1), w)
        call("tn_conv_weight_gemm", ptr(w), ptr(w3), Co, Ci, K, Kpad, 1)
        ctx.meta = (Co, Ci, K, Kpad)
        return w3

    @staticmethod
    def backward(ctx, dw3):
        Co, Ci, K, Kpad = ctx.meta
        dw = gempty((Co, Ci, K), dw3)
        call("tn_conv_weight_gemm", ptr(_c(dw3)), ptr(dw), Co, Ci, K, Kpad, 0)
        return dw, None


# TN_PROLOG_TC=0: keep the K-tap dense conv (prolog) on the exact-fp32 CUDA-core kernel (A/B)
PROLOG_TC = __import__("os").environ.get("TN_PROLOG_TC", "1") != "0"


def conv_ktap_as_gemm_ok(x: Tensor, w: Tensor, B: int, T: int) -> bool:
    """A dense K-tap conv whose unrolled reduction fits one GEMM tile row (Ci * K <= 512) and whose input needs no gradient."""
    Co, Ci, K = w.shape
    return (TC_ENABLED and PROLOG_TC and K > 1 and not x.requires_grad and B * T >= 512 and Co % 128 == 0 and Ci * K <= 512)


def _bn_trainable(bn: torch.nn.BatchNorm1d) -> bool:
    return (bn.training or bn.running_mean is None) and bn.momentum is not None and bn.affine


def conv_gemm_bn(x, w, bias, bn: torch.nn.BatchNorm1d, B: int, T: int):
    """(z, scale, shift) of ``bn(conv(x))``: fused in train mode, conv + running-statistics fold in eval mode."""
    if _bn_trainable(bn):
        return ConvGemmBN.apply(x, w, bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                                bn.momentum, bn.eps, B, T)
    z, stats = conv_gemm(x, w, bias, B, T, want_stats=bn.training)
    scale, shift = bn_fold(stats, bn, float(B * T))
    return z, scale, shift


class ColStats(Function):
    """stats = [sum_r x | sum_r x^2] per channel (fp64)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        R, C = x.shape
        stats = empty((2 * C,), x, torch.float64)
        call("tn_colstats", ptr(x), ptr(stats), R, C)
        ctx.save_for_backward(x)
        return stats

    @staticmethod
    def backward(ctx, dstats):
        (x,) = ctx.saved_tensors
        dx = empty(x.shape, x)
        call("tn_stats_bwd", None, ptr(x), ptr(dstats.contiguous()), ptr(dx), None, x.shape[0], x.shape[1])
        return dx


class BNFold(Function):
    """BatchNorm1d folded to per-channel (scale, shift).

    training: batch statistics from ``stats`` over ``n`` samples, biased variance; running
    statistics updated in place with ``momentum`` (unbiased variance) and
    ``num_batches_tracked += 1``.  eval: running statistics.  nn.BatchNorm1d semantics
    (reference: src/modules.py:128, src/models.py:454,506,512)."""

    @staticmethod
    def forward(ctx, stats, gamma, beta, running_mean, running_var, nbt, n: float, momentum: float, eps: float,
                training: bool):
        gamma, beta = _c(gamma), _c(beta)
        C = gamma.numel()
        scale, shift = empty((C,), gamma), empty((C,), gamma)
        mean, invstd = empty((C,), gamma), empty((C,), gamma)
        call("tn_bn_finalize", ptr(stats), float(n), ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), ptr(nbt),
             float(momentum), float(eps), 1 if training else 0, ptr(scale), ptr(shift), ptr(mean), ptr(invstd), C)
        ctx.save_for_backward(gamma, mean, invstd)
        ctx.meta = (float(n), training, stats is not None)
        return scale, shift

    @staticmethod
    def backward(ctx, dscale, dshift):
        gamma, mean, invstd = ctx.saved_tensors
        n, training, has_stats = ctx.meta
        C = gamma.numel()
        dscale = _c(dscale) if dscale is not None else zeros((C,), gamma)
        dshift = _c(dshift) if dshift is not None else zeros((C,), gamma)
        dgamma, dbeta = gempty((C,), gamma), gempty((C,), gamma)
        dstats = empty((2 * C,), gamma, torch.float64) if (training and has_stats) else None
        call("tn_bn_bwd_coef", ptr(dscale), ptr(dshift), ptr(mean), ptr(invstd), ptr(gamma), n, 1 if dstats is not None else 0,
             ptr(dgamma), ptr(dbeta), ptr(dstats), C)
        return dstats, dgamma, dbeta, None, None, None, None, None, None, None


def bn_fold(stats, bn: torch.nn.BatchNorm1d, n: float):
    """(scale, shift) of ``bn`` applied to a tensor whose statistics are ``stats``."""
    training = bn.training or (bn.running_mean is None)
    if bn.momentum is None:
        raise NotImplementedError("BatchNorm1d(momentum=None) (cumulative average) is not supported")
    return BNFold.apply(stats if training else None, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                        bn.num_batches_tracked if training else None, n, bn.momentum, bn.eps, training)


# ----------------------------------------------------------------------------
# lazy activation: materialise / depthwise conv
# ----------------------------------------------------------------------------
class Act(Function):
    """y = dropout(relu(z * scale + shift))."""

    @staticmethod
    def forward(ctx, z, scale, shift, seed, relu: bool, p: float, layer: int):
        z, scale, shift = _c(z), _c(scale), _c(shift)
        R, C = z.shape
        y = empty(z.shape, z)
        if p > 0 and seed is None:
            raise ValueError("dropout needs a seed tensor")
        call("tn_act_fwd", ptr(z), ptr(y), ptr(scale), ptr(shift), int(relu), float(p), ptr(seed), int(layer), R, C)
        ctx.save_for_backward(z, scale, shift, seed)
        ctx.meta = (relu, p, layer)
        return y

    @staticmethod
    def backward(ctx, dy):
        z, scale, shift, seed = ctx.saved_tensors
        relu, p, layer = ctx.meta
        R, C = z.shape
        dy = _c(dy)
        dz = empty(z.shape, z)
        dscale, dshift = zeros((C,), z), zeros((C,), z)
        call("tn_act_bwd", ptr(dy), ptr(z), ptr(dz), ptr(dscale), ptr(dshift), ptr(scale), ptr(shift), int(relu), float(p),
             ptr(seed), int(layer), R, C)
        return dz, dscale, dshift, None, None, None, None


class Act2(Function):
    """``Act`` for an activation with TWO consumers (the epilog output feeds the attention MLP and the pooling,
    src/models.py:564-584): returns the materialised tensor twice (two aliases of one buffer), so that each consumer's
    gradient arrives on its own and ``tn_act_bwd2`` adds them while loading -- autograd's sum of the two [R, C]
    gradients (read 2, write 1) never runs."""

    @staticmethod
    def forward(ctx, z, scale, shift, seed, relu: bool, p: float, layer: int):
        z, scale, shift = _c(z), _c(scale), _c(shift)
        R, C = z.shape
        y = empty(z.shape, z)
        if p > 0 and seed is None:
            raise ValueError("dropout needs a seed tensor")
        call("tn_act_fwd", ptr(z), ptr(y), ptr(scale), ptr(shift), int(relu), float(p), ptr(seed), int(layer), R, C)
        ctx.save_for_backward(z, scale, shift, seed)
        ctx.meta = (relu, p, layer)
        ctx.set_materialize_grads(False)
        return y, y.detach()

    @staticmethod
    def backward(ctx, dy1, dy2):
        z, scale, shift, seed = ctx.saved_tensors
        relu, p, layer = ctx.meta
        R, C = z.shape
        if dy1 is None:
            dy1, dy2 = dy2, None
        if dy1 is None:
            return None, None, None, None, None, None, None
        dy1 = _c(dy1)
        dy2 = _c(dy2) if dy2 is not None else None
        dz = empty(z.shape, z)
        dscale, dshift = zeros((C,), z), zeros((C,), z)
        call("tn_act_bwd2", ptr(dy1), ptr(dy2), ptr(z), ptr(dz), ptr(dscale), ptr(dshift), ptr(scale), ptr(shift), int(relu),
             float(p), ptr(seed), int(layer), R, C)
        return dz, dscale, dshift, None, None, None, None


class Depthwise(Function):
    """u = depthwise_K(act(z)) + bias; act = identity when scale is None."""

    @staticmethod
    def forward(ctx, z, scale, shift, w, bias, seed, relu: bool, p: float, layer: int, B: int, T: int):
        z, scale, shift, w, bias = _c(z), _c(scale), _c(shift), _c(w), _c(bias)
        C, K = w.shape[0], w.shape[-1]
        assert z.shape == (B * T, C)
        u = empty(z.shape, z)
        call("tn_dw_fwd", ptr(z), ptr(u), ptr(w), ptr(bias), ptr(scale), ptr(shift), int(relu), float(p), ptr(seed), int(layer),
             B, T, C, K)
        ctx.save_for_backward(z, scale, shift, w, bias, seed)
        ctx.meta = (relu, p, layer, B, T)
        return u

    @staticmethod
    def backward(ctx, du):
        z, scale, shift, w, bias, seed = ctx.saved_tensors
        relu, p, layer, B, T = ctx.meta
        C, K = w.shape[0], w.shape[-1]
        du = _c(du)
        dz = empty(z.shape, z)
        dw = zeros(w.shape, w)
        db = zeros((C,), z) if bias is not None else None
        dscale = zeros((C,), z) if scale is not None else None
        dshift = zeros((C,), z) if scale is not None else None
        call("tn_dw_bwd", ptr(du), ptr(z), ptr(dz), ptr(w), ptr(dw), ptr(db), ptr(dscale), ptr(dshift), ptr(scale), ptr(shift),
             int(relu), float(p), ptr(seed), int(layer), B, T, C, K)
        return dz, dscale, dshift, dw, db, None, None, None, None, None, None


TC_FUSE_DWBWD = __import__("os").environ.get("TN_FUSE_DWBWD", "1") != "0"     # 0: tn_gemm_tc (dgrad) + tn_dw_bwd, for A/B
# Two launch fusions, measured on the graph-replayed TitaNet-S step and OFF by default because they lose:
#   TN_FUSE_BLOCK_ENTRY=1  first sub-block + skip branch as one autograd node, the skip data gradient accumulated in the GEMM
#                          epilogue (red.global.add) instead of an add kernel:   10.21 -> 10.24 ms
#   TN_FUSE_SE_MLP=1       SE MLP backward run by the last block of tail_bwd1 (-17 launches; with the forward twin, removed in
#                          round 2 because its mean used float atomics: 10.21 -> 10.32 ms)
# (inside a CUDA graph a launch boundary costs less than the serial tail the fused kernels add)
import os as _os
FUSE_BLOCK_ENTRY = _os.environ.get("TN_FUSE_BLOCK_ENTRY", "0") == "1"
FUSE_SE_MLP = _os.environ.get("TN_FUSE_SE_MLP", "0") == "1"
# squeeze + excitation forward as one cluster kernel (tn_se_squeeze_excite: DSMEM exchange, no atomics); 0 = two launches (A/B)
FUSE_SE_FWD = _os.environ.get("TN_FUSE_SE_FWD", "1") != "0"
# ... and the mega-block tail in the same launch (tn_se_tail_fwd: the activated z3 tile stays in shared memory); 0 = separate tail (A/B)
FUSE_SE_TAIL = _os.environ.get("TN_FUSE_SE_TAIL", "1") != "0"
# the block output as two aliases (ops.SETail2): the next block's two consumers send their gradients separately and
# tn_tail_bwd1s adds them on load; 0 = one output and autograd's sum kernel (A/B)
SE_TAIL_TWO_OUT = _os.environ.get("TN_SE_TAIL_TWO_OUT", "1") != "0"


# TN_FUSE_DWFWD=1: the depthwise conv as the pointwise GEMM's operand producer (tn_gemm_tc_dwfwd).  Parity-green and 1.3 us
# faster per sub-block than tn_dw_fwd + GEMM in isolation (33.4 vs 34.7 us), but transform-bound, and on the whole graph-replayed
# step it measures slower (10.18 vs 10.00 ms, same box), so it is opt-in until the producer keeps up with the tensor pipe.
TC_FUSE_DWFWD = _os.environ.get("TN_FUSE_DWFWD", "0") == "1"


def _dw_pw_forward(z, scale, shift, dw_w, dw_b, pw_w, pw_b, seed, relu, p, layer, B, T, stats, bn):
    """u = depthwise(act(z)) + b_dw ; zo = pointwise(u) + b_pw (+ statistics / BatchNorm fold of zo).  One tensor-core kernel
    (tn_gemm_tc_dwfwd: the depthwise conv is the GEMM's operand producer) when the shape allows, else tn_dw_fwd + GEMM.
    Returns (u, zo, ws_dgrad)."""
    C, K = dw_w.shape[0], dw_w.shape[-1]
    pw3 = pw_w if pw_w.dim() == 3 else pw_w.unsqueeze(-1)
    Co, R = pw3.shape[0], B * T
    u = empty(z.shape, z)
    zo = empty((R, Co), z)
    sp = cached_splits(pw_w)
    if (TC_FUSE_DWFWD and TC_FWD_NSPLIT == 3 and _tc_ok(R, C, Co, pw3.shape[2]) and K % 2 == 1 and K <= 7 and (R + 16) * C < 2 ** 32):
        ws = sp[0] if sp else None
        if ws is None:
            ws = torch.empty((WS_PLANES, Co, C), device=z.device, dtype=torch.float32)
            call("tn_split_tf32", ptr(pw3), ptr(ws), Co, C, 0)
        sc, keep = scratch(z) if stats is not None else (None, None)
        call("tn_gemm_tc_dwfwd", ptr(z), ptr(ws), ptr(dw_w), ptr(dw_b), ptr(scale), ptr(shift), int(relu), float(p), ptr(seed),
             int(layer), ptr(pw_b), ptr(u), ptr(zo), ptr(stats), ctypes.byref(bn) if bn is not None else None, B, T, C, Co, K,
             ctypes.byref(sc) if sc is not None else None, tag=f"dw+fwd R{R} Ci{C} Co{Co} K1")
    else:
        call("tn_dw_fwd", ptr(z), ptr(u), ptr(dw_w), ptr(dw_b), ptr(scale), ptr(shift), int(relu), float(p), ptr(seed), int(layer),
             B, T, C, K)
        _gemm_fwd(u, pw3, pw_b, zo, stats, B, T, 0, 0, ws=sp[0] if sp else None, bn=bn)
    return u, zo, (sp[1] if sp else None)


class DwPw(Function):
    """Depthwise-separable convolution on a lazy activation, with the following BatchNorm's
    statistics:  z_out = pointwise(depthwise_K(act(z)) + b_dw) + b_pw  (modules.DepthwiseConv1d,
    src/modules.py:64-79).  Backward runs the pointwise data-gradient GEMM and the whole depthwise /
    BN-ReLU-dropout backward of the previous layer as ONE tensor-core kernel (tn_gemm_tc_dwbwd)."""

    @staticmethod
    def forward(ctx, z, scale, shift, dw_w, dw_b, pw_w, pw_b, seed, relu: bool, p: float, layer: int, B: int, T: int,
                want_stats: bool):
        z, scale, shift, dw_w, dw_b, pw_w, pw_b = map(_c, (z, scale, shift, dw_w, dw_b, pw_w, pw_b))
        C, K = dw_w.shape[0], dw_w.shape[-1]
        Co = pw_w.shape[0]
        assert z.shape == (B * T, C)
        stats = empty((2 * Co,), z, torch.float64) if want_stats else None
        u, zo, ctx.ws_t = _dw_pw_forward(z, scale, shift, dw_w, dw_b, pw_w, pw_b, seed, relu, p, layer, B, T, stats, None)
        ctx.save_for_backward(z, scale, shift, dw_w, dw_b, pw_w, pw_b, seed, u, zo if want_stats else None)
        ctx.meta = (relu, p, layer, B, T, want_stats)
        return zo, stats

    @staticmethod
    def backward(ctx, dzo, dstats):
        z, scale, shift, dw_w, dw_b, pw_w, pw_b, seed, u, zo = ctx.saved_tensors
        relu, p, layer, B, T, want_stats = ctx.meta
        C, K = dw_w.shape[0], dw_w.shape[-1]
        pw3 = pw_w if pw_w.dim() == 3 else pw_w.unsqueeze(-1)
        Co, R = pw3.shape[0], B * T
        dz = _c(dzo) if dzo is not None else zeros((R, Co), z)
        db_pw = zeros((Co,), z) if pw_b is not None else None
        db_done = False
        if want_stats and dstats is not None:
            g = empty(dz.shape, dz)
            call("tn_stats_bwd", ptr(dz), ptr(zo), ptr(dstats.contiguous()), ptr(g), ptr(db_pw), R, Co)
            dz, db_done = g, True
        dpw = zeros(pw_w.shape, pw_w)
        _gemm_wgrad(dz, u, dpw if dpw.dim() == 3 else dpw.unsqueeze(-1), None if db_done else db_pw, B, T)
        dzp, dscale, dshift, ddw, ddb = _dwpw_dgrad(dz, pw3, ctx.ws_t, z, scale, shift, dw_w, dw_b, seed, relu, p, layer, B, T)
        return dzp, dscale, dshift, ddw, ddb, dpw, db_pw, None, None, None, None, None, None, None


def _dwpw_dgrad(dz, pw3, ws_t, z, scale, shift, dw_w, dw_b, seed, relu, p, layer, B, T, bnb=None):
    """Data gradient of pointwise(depthwise(act(z))): one fused tensor-core kernel when the shape allows.  ``bnb`` (TnBnBwd):
    ``dz`` is the direct gradient w.r.t. the block's pre-BatchNorm output and the BatchNorm backward runs inside the kernel."""
    C, K = dw_w.shape[0], dw_w.shape[-1]
    Co, R = pw3.shape[0], B * T
    dzp = empty(z.shape, z)
    ddw = zeros(dw_w.shape, dw_w)
    ddb = zeros((C,), z) if dw_b is not None else None
    dscale = zeros((C,), z) if scale is not None else None
    dshift = zeros((C,), z) if scale is not None else None
    if _dwbwd_tc_ok(R, Co, C, K):
        ws = ws_t
        if ws is None:
            ws = torch.empty((WS_PLANES, C, Co), device=z.device, dtype=torch.float32)
            call("tn_split_tf32", ptr(pw3), ptr(ws), C, Co, 1)
        if bnb is not None:
            call("tn_gemm_tc_dwbwd_bn", ptr(dz), ptr(ws), ctypes.byref(bnb), ptr(z), ptr(dzp), ptr(dw_w), ptr(ddw), ptr(ddb), ptr(dscale),
                 ptr(dshift), ptr(scale), ptr(shift), int(relu), float(p), ptr(seed), int(layer), B, T, Co, C, K,
                 tag=f"bnbwd+dgrad+dwbwd R{R} Ci{Co} Co{C} K1")
        else:
            call("tn_gemm_tc_dwbwd", ptr(dz), ptr(ws), ptr(z), ptr(dzp), ptr(dw_w), ptr(ddw), ptr(ddb), ptr(dscale), ptr(dshift),
                 ptr(scale), ptr(shift), int(relu), float(p), ptr(seed), int(layer), B, T, Co, C, K, TC_BWD_NSPLIT,
                 tag=f"dgrad+dwbwd R{R} Ci{Co} Co{C} K1")
    else:
        assert bnb is None, "the fused BatchNorm backward needs the tensor-core depthwise-backward kernel"
        du = empty(z.shape, z)
        _gemm_fwd(dz, pw3, None, du, None, B, T, 1, 0, ws=ws_t)
        call("tn_dw_bwd", ptr(du), ptr(z), ptr(dzp), ptr(dw_w), ptr(ddw), ptr(ddb), ptr(dscale), ptr(dshift), ptr(scale),
             ptr(shift), int(relu), float(p), ptr(seed), int(layer), B, T, C, K)
    return dzp, dscale, dshift, ddw, ddb


# TN_RECOMPUTE_U=1: do not keep the depthwise outputs u for the backward pass (3 of the 9 tensors a mega-block saves); the
# weight-gradient GEMM's operand is recomputed by one tn_dw_fwd launch per sub-block (same dropout seed and layer id: the same
# values).  ~6 % slower per step, one third less activation memory: what lets the batch sweep reach 2048 on one GPU.
RECOMPUTE_U = __import__("os").environ.get("TN_RECOMPUTE_U", "0") == "1"


def _dwbwd_tc_ok(R: int, Co: int, C: int, K: int) -> bool:
    return (TC_ENABLED and TC_FUSE_DWBWD and R >= TC_MIN_ROWS and Co % 32 == 0 and C % 128 == 0 and K % 2 == 1 and K <= 11
            and (R + 16) * C < 2 ** 32)


class DwPwBN(Function):
    """``DwPw`` followed by a TRAIN-mode BatchNorm1d folded by the GEMM's last CTA (see ``ConvGemmBN``):
    returns (z_out, scale_out, shift_out).  One ConvBlock1d(depthwise=True) of a mega-block
    (src/modules.py:119-134) = tn_dw_fwd + tn_gemm_tc_bn forward, tn_bn_stats_bwd + tn_wgrad_tc +
    tn_gemm_tc_dwbwd backward."""

    @staticmethod
    def forward(ctx, z, scale, shift, dw_w, dw_b, pw_w, pw_b, gamma, beta, rm, rv, nbt, momentum: float, eps: float, seed,
                relu: bool, p: float, layer: int, B: int, T: int):
        z, scale, shift, dw_w, dw_b, pw_w, pw_b, gamma, beta = map(_c, (z, scale, shift, dw_w, dw_b, pw_w, pw_b, gamma, beta))
        C, K = dw_w.shape[0], dw_w.shape[-1]
        Co = pw_w.shape[0]
        assert z.shape == (B * T, C)
        stats, fold = _bn_forward_buffers(Co, z)
        n = float(B * T)
        bn = make_bn_fold(gamma, beta, rm, rv, nbt, momentum, eps, n, fold[0], fold[1], fold[2], fold[3])
        u, zo, ctx.ws_t = _dw_pw_forward(z, scale, shift, dw_w, dw_b, pw_w, pw_b, seed, relu, p, layer, B, T, stats, bn)
        ctx.save_for_backward(z, scale, shift, dw_w, dw_b, pw_w, pw_b, seed, None if RECOMPUTE_U else u, zo, gamma, fold)
        ctx.meta = (relu, p, layer, B, T, n)
        return zo, fold[0], fold[1]

    @staticmethod
    def backward(ctx, dzo, dscale_o, dshift_o):
        z, scale, shift, dw_w, dw_b, pw_w, pw_b, seed, u, zo, gamma, fold = ctx.saved_tensors
        relu, p, layer, B, T, n = ctx.meta
        pw3 = pw_w if pw_w.dim() == 3 else pw_w.unsqueeze(-1)
        Co = pw3.shape[0]
        if u is None:                 # TN_RECOMPUTE_U: rebuild the wgrad operand (identical values: stateless dropout hash)
            u = empty(z.shape, z)
            call("tn_dw_fwd", ptr(z), ptr(u), ptr(dw_w), ptr(dw_b), ptr(scale), ptr(shift), int(relu), float(p), ptr(seed), int(layer),
                 B, T, dw_w.shape[0], dw_w.shape[-1])
        R, C, K = B * T, dw_w.shape[0], dw_w.shape[-1]
        if _bnbwd_fusable(dzo, R, Co, C) and _dwbwd_tc_ok(R, Co, C, K):
            # data gradient first: its operand producer computes g (the BatchNorm backward) and writes it for the wgrad below
            bnb, g, db_pw, dgamma, dbeta, keep = _make_bn_bwd(_c(dzo), zo, dscale_o, dshift_o, fold, gamma, n, pw_b, Co)
            dzp, dscale, dshift, ddw, ddb = _dwpw_dgrad(_c(dzo), pw3, ctx.ws_t, z, scale, shift, dw_w, dw_b, seed, relu, p, layer, B, T,
                                                        bnb=bnb)
            dpw = zeros(pw_w.shape, pw_w)
            _gemm_wgrad(g, u, dpw if dpw.dim() == 3 else dpw.unsqueeze(-1), None, B, T)
        else:
            g, db_pw, dgamma, dbeta = _bn_backward(dzo, zo, dscale_o, dshift_o, fold, gamma, n, pw_b, R, Co)
            dpw = zeros(pw_w.shape, pw_w)
            _gemm_wgrad(g, u, dpw if dpw.dim() == 3 else dpw.unsqueeze(-1), None, B, T)
            dzp, dscale, dshift, ddw, ddb = _dwpw_dgrad(g, pw3, ctx.ws_t, z, scale, shift, dw_w, dw_b, seed, relu, p, layer, B, T)
        return (dzp, dscale, dshift, ddw, ddb, dpw, db_pw, dgamma, dbeta, None, None, None, None, None, None, None, None,
                None, None, None)


class BlockEntryBN(Function):
    """The two consumers of a mega-block's input in ONE autograd node (train mode): the first sub-block
    (depthwise-separable conv + BatchNorm, ``DwPwBN``) and the skip branch (1x1 conv + BatchNorm, ``ConvGemmBN``)
    (src/models.py:435-455, 467-469).  Forward is the same kernels; in backward the skip branch's data-gradient GEMM
    accumulates straight into the gradient the fused depthwise backward wrote (TN_EPI_ACCUM), so autograd has no
    ``dx_a + dx_b`` kernel to run on the 17 block inputs.  ``z`` is a plain activation (the previous block's output)."""

    @staticmethod
    def forward(ctx, z, dw_w, dw_b, pw_w, pw_b, g1, b1, rm1, rv1, nbt1, mom1: float, eps1: float,
                sk_w, sk_b, gs, bs, rms, rvs, nbts, moms: float, epss: float, B: int, T: int):
        z, dw_w, dw_b, pw_w, pw_b, g1, b1, sk_w, sk_b, gs, bs = map(_c, (z, dw_w, dw_b, pw_w, pw_b, g1, b1, sk_w, sk_b, gs, bs))
        C, K = dw_w.shape[0], dw_w.shape[-1]
        Co, Cs = pw_w.shape[0], sk_w.shape[0]
        R = B * T
        assert z.shape == (R, C)
        n = float(R)
        # skip branch first, like the reference's forward (src/models.py:469)
        s = empty((R, Cs), z)
        st_s, fold_s = _bn_forward_buffers(Cs, z)
        bn_s = make_bn_fold(gs, bs, rms, rvs, nbts, moms, epss, n, fold_s[0], fold_s[1], fold_s[2], fold_s[3])
        sp_s = cached_splits(sk_w)
        sk3 = sk_w if sk_w.dim() == 3 else sk_w.unsqueeze(-1)
        _gemm_fwd(z, sk3, sk_b, s, st_s, B, T, 0, 0, ws=sp_s[0] if sp_s else None, bn=bn_s)
        st_1, fold_1 = _bn_forward_buffers(Co, z)
        bn_1 = make_bn_fold(g1, b1, rm1, rv1, nbt1, mom1, eps1, n, fold_1[0], fold_1[1], fold_1[2], fold_1[3])
        u, zo, ctx.ws_t1 = _dw_pw_forward(z, None, None, dw_w, dw_b, pw_w, pw_b, None, False, 0.0, 0, B, T, st_1, bn_1)
        ctx.ws_ts = sp_s[1] if sp_s else None
        ctx.save_for_backward(z, dw_w, dw_b, pw_w, pw_b, g1, sk_w, sk_b, gs, u, zo, s, fold_1, fold_s)
        ctx.meta = (B, T, n)
        return zo, fold_1[0], fold_1[1], s, fold_s[0], fold_s[1]

    @staticmethod
    def backward(ctx, dzo, dsc1, dsh1, ds, dscs, dshs):
        z, dw_w, dw_b, pw_w, pw_b, g1, sk_w, sk_b, gs, u, zo, s, fold_1, fold_s = ctx.saved_tensors
        B, T, n = ctx.meta
        R = B * T
        pw3 = pw_w if pw_w.dim() == 3 else pw_w.unsqueeze(-1)
        sk3 = sk_w if sk_w.dim() == 3 else sk_w.unsqueeze(-1)
        Co, Cs = pw3.shape[0], sk3.shape[0]
        ga, db_pw, dg1, db1 = _bn_backward(dzo, zo, dsc1, dsh1, fold_1, g1, n, pw_b, R, Co)
        dpw = zeros(pw_w.shape, pw_w)
        _gemm_wgrad(ga, u, dpw if dpw.dim() == 3 else dpw.unsqueeze(-1), None, B, T)
        dz, _, _, ddw, ddb = _dwpw_dgrad(ga, pw3, ctx.ws_t1, z, None, None, dw_w, dw_b, None, False, 0.0, 0, B, T)
        gb, db_s, dgs, dbs = _bn_backward(ds, s, dscs, dshs, fold_s, gs, n, sk_b, R, Cs)
        dws = zeros(sk_w.shape, sk_w)
        _gemm_wgrad(gb, z, dws if dws.dim() == 3 else dws.unsqueeze(-1), None, B, T)
        _gemm_fwd(gb, sk3, None, dz, None, B, T, 1, EPI_ACCUM, ws=ctx.ws_ts)          # dz += gb W_skip
        return (dz, ddw, ddb, dpw, db_pw, dg1, db1, None, None, None, None, None,
                dws, db_s, dgs, dbs, None, None, None, None, None, None, None)


# ----------------------------------------------------------------------------
# squeeze-excitation + mega-block tail
# ----------------------------------------------------------------------------
class SETail(Function):
    """out = dropout(relu( (s*scale_s+shift_s) + gate * a3 )),  gate = SE(mean_t a3),
    a3 = dropout(relu(z3*scale3+shift3)).  (src/modules.py:173-189, src/models.py:467-472)"""

    @staticmethod
    def forward(ctx, z3, sc3, sh3, s, scs, shs, W1, W2, seed, p3: float, layer3: int, p_o: float, layer_o: int, B: int,
                T: int):
        z3, sc3, sh3, s, scs, shs, W1, W2 = map(_c, (z3, sc3, sh3, s, scs, shs, W1, W2))
        C = z3.shape[1]
        Cr = W1.shape[0]
        m, gate = empty((B, C), z3), empty((B, C), z3)
        out = empty(z3.shape, z3)
        if FUSE_SE_TAIL and LIB.query("tn_se_tail_fwd_supported", T, C, Cr):
            # squeeze + excitation + tail in one cluster kernel: the activated z3 tile stays in shared memory in between
            call("tn_se_tail_fwd", ptr(z3), ptr(s), ptr(m), ptr(gate), ptr(out), ptr(W1), ptr(W2), ptr(sc3), ptr(sh3), float(p3),
                 int(layer3), ptr(scs), ptr(shs), float(p_o), int(layer_o), ptr(seed), B, T, C, Cr)
            ctx.save_for_backward(z3, sc3, sh3, s, scs, shs, W1, W2, seed, m, gate, out)
            ctx.meta = (p3, layer3, p_o, layer_o, B, T)
            return out
        if FUSE_SE_FWD and LIB.query("tn_se_squeeze_excite_supported", C, Cr):
            call("tn_se_squeeze_excite", ptr(z3), ptr(m), ptr(gate), ptr(W1), ptr(W2), ptr(sc3), ptr(sh3), 1, float(p3), ptr(seed),
                 int(layer3), B, T, C, Cr)
        else:
            call("tn_se_mean", ptr(z3), ptr(m), ptr(sc3), ptr(sh3), 1, float(p3), ptr(seed), int(layer3), B, T, C)
            call("tn_se_mlp_fwd", ptr(m), ptr(W1), ptr(W2), ptr(gate), B, C, Cr)
        call("tn_tail_fwd", ptr(z3), ptr(s), ptr(gate), ptr(out), ptr(sc3), ptr(sh3), float(p3), int(layer3), ptr(scs), ptr(shs),
             float(p_o), int(layer_o), ptr(seed), B, T, C)
        ctx.save_for_backward(z3, sc3, sh3, s, scs, shs, W1, W2, seed, m, gate, out)
        ctx.meta = (p3, layer3, p_o, layer_o, B, T)
        return out

    @staticmethod
    def backward(ctx, dout):
        return _se_tail_backward(ctx, dout, None)


def _se_tail_backward(ctx, dout, dout2):
    """Backward of SETail / SETail2.  ``dout2``: the gradient that reached the second alias of the block output (or None);
    tn_tail_bwd1s adds the two while loading and writes the sum once for pass 2."""
    z3, sc3, sh3, s, scs, shs, W1, W2, seed, m, gate, out = ctx.saved_tensors
    p3, layer3, p_o, layer_o, B, T = ctx.meta
    C, Cr = z3.shape[1], W1.shape[0]
    if dout is None:
        dout, dout2 = dout2, None
    if dout is None:
        return (None,) * 15
    dout = _c(dout)
    dgate = zeros((B, C), z3)
    dm = empty((B, C), z3)
    dW1, dW2 = zeros(W1.shape, W1), zeros(W2.shape, W2)
    if dout2 is not None:
        dsum = empty(z3.shape, z3)
        call("tn_tail_bwd1s", ptr(dout), ptr(_c(dout2)), ptr(dsum), ptr(out), ptr(z3), ptr(dgate), ptr(sc3), ptr(sh3), float(p3),
             int(layer3), float(p_o), ptr(seed), B, T, C)
        call("tn_se_mlp_bwd", ptr(dgate), ptr(gate), ptr(m), ptr(W1), ptr(W2), ptr(dm), ptr(dW1), ptr(dW2), B, C, Cr)
        dout = dsum
    elif FUSE_SE_MLP:
        tickets = zeros((B,), z3, torch.int32)
        call("tn_tail_bwd1_mlp", ptr(dout), ptr(out), ptr(z3), ptr(dgate), ptr(tickets), ptr(gate), ptr(m), ptr(W1), ptr(W2), ptr(dm),
             ptr(dW1), ptr(dW2), ptr(sc3), ptr(sh3), float(p3), int(layer3), float(p_o), ptr(seed), B, T, C, Cr)
    else:
        call("tn_tail_bwd1", ptr(dout), ptr(out), ptr(z3), ptr(dgate), ptr(sc3), ptr(sh3), float(p3), int(layer3), float(p_o),
             ptr(seed), B, T, C)
        call("tn_se_mlp_bwd", ptr(dgate), ptr(gate), ptr(m), ptr(W1), ptr(W2), ptr(dm), ptr(dW1), ptr(dW2), B, C, Cr)
    dz3, ds = empty(z3.shape, z3), empty(z3.shape, z3)
    red = zeros((4, C), z3)
    call("tn_tail_bwd2", ptr(dout), ptr(out), ptr(z3), ptr(s), ptr(gate), ptr(dm), ptr(dz3), ptr(ds), red[0].data_ptr(),
         red[1].data_ptr(), red[2].data_ptr(), red[3].data_ptr(), ptr(sc3), ptr(sh3), float(p3), int(layer3), ptr(scs),
         ptr(shs), float(p_o), ptr(seed), B, T, C)
    return dz3, red[0], red[1], ds, red[2], red[3], dW1, dW2, None, None, None, None, None, None, None


class SETail2(Function):
    """``SETail`` whose output has TWO consumers (the next mega-block's skip conv and its first depthwise conv,
    src/models.py:467-469): returns the block output twice (two aliases of one buffer) so that each consumer's gradient
    arrives on its own and the backward adds them while loading (tn_tail_bwd1s) -- autograd's sum kernel on the 16 inner block
    outputs (read 2, write 1 [B*T, C] tensors each) never runs."""

    @staticmethod
    def forward(ctx, *args):
        out = SETail.forward(ctx, *args)
        ctx.set_materialize_grads(False)
        return out, out.detach()

    @staticmethod
    def backward(ctx, dout, dout2):
        return _se_tail_backward(ctx, dout, dout2)


class MeanT(Function):
    """m[B, C] = mean over time of a plain [B*T, C] tensor (nn.AdaptiveAvgPool1d(1): src/modules.py:165,
    src/models.py:498); backward broadcasts dm / T over the frames."""

    @staticmethod
    def forward(ctx, x, B: int, T: int):
        x = _c(x)
        C = x.shape[1]
        m = empty((B, C), x)
        call("tn_se_mean", ptr(x), ptr(m), None, None, 0, 0.0, None, 0, B, T, C)
        ctx.meta = (B, T, C)
        return m

    @staticmethod
    def backward(ctx, dm):
        B, T, C = ctx.meta
        dm = _c(dm)
        dx = empty((B * T, C), dm)
        call("tn_bcast_rows", ptr(dm), ptr(dx), 1.0 / T, B, T, C)
        return dx, None, None


class SEMlp(Function):
    """gate = sigmoid(W2 relu(W1 m)), no biases (SqueezeExcitation.excitation, src/modules.py:166-171)."""

    @staticmethod
    def forward(ctx, m, W1, W2):
        m, W1, W2 = _c(m), _c(W1), _c(W2)
        B, C = m.shape
        gate = empty((B, C), m)
        call("tn_se_mlp_fwd", ptr(m), ptr(W1), ptr(W2), ptr(gate), B, C, W1.shape[0])
        ctx.save_for_backward(m, W1, W2, gate)
        return gate

    @staticmethod
    def backward(ctx, dgate):
        m, W1, W2, gate = ctx.saved_tensors
        B, C = m.shape
        dm = empty((B, C), m)
        dW1, dW2 = zeros(W1.shape, W1), zeros(W2.shape, W2)
        call("tn_se_mlp_bwd", ptr(_c(dgate)), ptr(gate), ptr(m), ptr(W1), ptr(W2), ptr(dm), ptr(dW1), ptr(dW2), B, C, W1.shape[0])
        return dm, dW1, dW2


class GateMul(Function):
    """out[b, t, c] = x[b, t, c] * gate[b, c]  (src/modules.py:187-189)."""

    @staticmethod
    def forward(ctx, x, gate, B: int, T: int):
        x, gate = _c(x), _c(gate)
        C = x.shape[1]
        out = empty(x.shape, x)
        call("tn_gate_mul_fwd", ptr(x), ptr(gate), ptr(out), B, T, C)
        ctx.save_for_backward(x, gate)
        ctx.meta = (B, T, C)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, gate = ctx.saved_tensors
        B, T, C = ctx.meta
        dx = empty(x.shape, x)
        dgate = zeros((B, C), x)
        call("tn_gate_mul_bwd", ptr(_c(dout)), ptr(x), ptr(gate), ptr(dx), ptr(dgate), B, T, C)
        return dx, dgate, None, None


# ----------------------------------------------------------------------------
# attentive statistics pooling
# ----------------------------------------------------------------------------
class ASPPool(Function):
    """pooled[B, 2D] = [sum_t alpha x | sqrt(clamp(sum_t alpha x^2 - mu^2, eps))],
    alpha = softmax_t(e).  (src/models.py:570-584)"""

    @staticmethod
    def forward(ctx, e, x, B: int, T: int, eps: float):
        e, x = _c(e), _c(x)
        D = e.shape[1]
        pooled, aux = empty((B, 2 * D), e), empty((B, 2, D), e)
        call("tn_asp_pool_fwd", ptr(e), ptr(x), ptr(pooled), ptr(aux), B, T, D, float(eps))
        ctx.save_for_backward(e, x, pooled, aux)
        ctx.meta = (B, T, eps)
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        e, x, pooled, aux = ctx.saved_tensors
        B, T, eps = ctx.meta
        D = e.shape[1]
        de, dx = empty(e.shape, e), empty(x.shape, x)
        call("tn_asp_pool_bwd", ptr(_c(dpooled)), ptr(pooled), ptr(aux), ptr(e), ptr(x), ptr(de), ptr(dx), B, T, D, float(eps))
        return de, dx, None, None, None


# ----------------------------------------------------------------------------
# embedding normalisation and loss heads
# ----------------------------------------------------------------------------
class L2Norm(Function):
    """y = x / max(||x||, eps) row-wise (eps = 0: plain division); also returns the norms."""

    @staticmethod
    def forward(ctx, x, eps: float):
        x = _c(x)
        B, E = x.shape
        y, norms = empty(x.shape, x), empty((B,), x)
        call("tn_l2norm_fwd", ptr(x), ptr(y), ptr(norms), B, E, float(eps))
        ctx.save_for_backward(y, norms)
        ctx.eps = eps
        return y, norms

    @staticmethod
    def backward(ctx, dy, dnorms):
        y, norms = ctx.saved_tensors
        B, E = y.shape
        dy = _c(dy) if dy is not None else zeros(y.shape, y)
        dx = empty(y.shape, y)
        call("tn_l2norm_bwd", ptr(dy), ptr(y), ptr(norms), ptr(_c(dnorms)) if dnorms is not None else None, ptr(dx), B, E,
             float(ctx.eps))
        return dx, None


class CrossEntropy(Function):
    """mean softmax cross-entropy + argmax (F.cross_entropy / torch.argmax; src/losses.py:40-42)."""

    @staticmethod
    def forward(ctx, logits, targets):
        logits = _c(logits)
        require_cuda(targets)
        targets = targets.to(torch.int64).contiguous()
        B, Cn = logits.shape
        loss_row, loss = empty((B,), logits), empty((), logits)
        preds = empty((B,), logits, torch.int64)
        call("tn_ce_fwd_bwd", ptr(logits), ptr(targets), ptr(loss_row), ptr(loss), ptr(preds), None, None, B, Cn)
        ctx.save_for_backward(logits, targets)
        ctx.mark_non_differentiable(preds)
        return loss, preds

    @staticmethod
    def backward(ctx, dloss, _dpreds):
        logits, targets = ctx.saved_tensors
        B, Cn = logits.shape
        loss_row, loss = empty((B,), logits), empty((), logits)
        preds = empty((B,), logits, torch.int64)
        dlogits = empty(logits.shape, logits)
        call("tn_ce_fwd_bwd", ptr(logits), ptr(targets), ptr(loss_row), ptr(loss), ptr(preds), ptr(dlogits), ptr(_c(dloss)), B, Cn)
        return dlogits, None


class AngularMargin(Function):
    """loss = -mean(num - log(exp(num) + sum_{j != y} exp(s c_j) + eps)),
    num = s (cos(m1 acos(c_y) + m2) - m3), c = clamp(raw, -1, 1)  (src/losses.py:101-130)."""

    @staticmethod
    def forward(ctx, raw, norms, targets, scale, m1: float, m2: float, m3: float, eps: float):
        raw = _c(raw)
        norms = _c(norms)
        require_cuda(targets)
        targets = targets.to(torch.int64).contiguous()
        B, Cn = raw.shape
        loss_row, loss = empty((B,), raw), empty((), raw)
        preds = empty((B,), raw, torch.int64)
        use_norm = scale is None
        call("tn_margin_fwd_bwd", ptr(raw), ptr(norms), ptr(targets), ptr(loss_row), ptr(loss), ptr(preds), None, None, None, B, Cn,
             0.0 if use_norm else float(scale), int(use_norm), float(m1), float(m2), float(m3), float(eps))
        ctx.save_for_backward(raw, norms, targets)
        ctx.meta = (scale, m1, m2, m3, eps)
        ctx.mark_non_differentiable(preds)
        return loss, preds

    @staticmethod
    def backward(ctx, dloss, _dpreds):
        raw, norms, targets = ctx.saved_tensors
        scale, m1, m2, m3, eps = ctx.meta
        B, Cn = raw.shape
        use_norm = scale is None
        loss_row, loss = empty((B,), raw), empty((), raw)
        preds = empty((B,), raw, torch.int64)
        draw = empty(raw.shape, raw)
        dnorm = empty((B,), raw) if use_norm else None
        call("tn_margin_fwd_bwd", ptr(raw), ptr(norms), ptr(targets), ptr(loss_row), ptr(loss), ptr(preds), ptr(draw), ptr(dnorm),
             ptr(_c(dloss)), B, Cn, 0.0 if use_norm else float(scale), int(use_norm), float(m1), float(m2), float(m3), float(eps))
        return draw, dnorm, None, None, None, None, None, None


def rownorm_(w: Tensor, eps: float = 1e-12) -> Tensor:
    """In-place F.normalize(w, dim=1) outside autograd (src/losses.py:86)."""
    require_cuda(w)
    assert w.is_contiguous() and w.dtype == torch.float32
    call("tn_rownorm_inplace", ptr(w), w.shape[0], w.shape[1], float(eps))
    torch.autograd.graph.increment_version(w)      # modified outside autograd's view (like ``.data = ...`` in the reference)
    return w


# ----------------------------------------------------------------------------
# mel front end (no gradient)
# ----------------------------------------------------------------------------
def mel_forward(wave: Tensor, lengths: Optional[Tensor], window: Tensor, fb: Tensor, band_lo: Tensor, band_hi: Tensor,
                n_fft: int, hop: int, n_mels: int, T_out: Optional[int] = None, nwc: bool = False,
                rates: Optional[Tensor] = None, frames: Optional[Tensor] = None, masks: Optional[Tensor] = None,
                n_fmask: int = 0, n_tmask: int = 0) -> Tensor:
    """wave [B, L] -> [B, n_mels, T] (or [B, T, n_mels] when nwc).  rates / frames / masks: SpecAugment draws
    (``tn_mel_specaug_fwd``: fp64 [B], int32 [B], int32 [B, n_fmask + n_tmask, 2])."""
    wave = _c(wave)
    B, L = wave.shape
    if T_out is None:
        T_out = 1 + L // hop
    if lengths is not None:
        require_cuda(lengths)
        lengths = lengths.to(torch.int32).contiguous()
    out = empty((B, T_out, n_mels) if nwc else (B, n_mels, T_out), wave)
    if rates is None and masks is None:
        call("tn_mel_fwd", ptr(wave), ptr(lengths), ptr(window), ptr(fb), ptr(band_lo), ptr(band_hi), ptr(out), B, L, L, T_out,
             n_fft, hop, n_mels, int(nwc))
        return out
    require_cuda(rates, frames, masks)
    if rates is not None and (rates.dtype != torch.float64 or frames.dtype != torch.int32 or rates.numel() != B or frames.numel() != B):
        raise TypeError("rates must be fp64 [B] and frames int32 [B]")
    if masks is not None and (masks.dtype != torch.int32 or tuple(masks.shape) != (B, n_fmask + n_tmask, 2)):
        raise TypeError("masks must be int32 [B, n_fmask + n_tmask, 2]")
    call("tn_mel_specaug_fwd", ptr(wave), ptr(lengths), ptr(window), ptr(fb), ptr(band_lo), ptr(band_hi),
         ptr(rates.contiguous() if rates is not None else None), ptr(frames.contiguous() if frames is not None else None),
         ptr(masks.contiguous() if masks is not None else None), n_fmask, n_tmask, ptr(out), B, L, L, T_out, n_fft, hop,
         n_mels, int(nwc))
    return out
