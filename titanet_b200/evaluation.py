"""Batched speaker-verification evaluation on the device.

Mirrors the evaluation half of the reference with the same names and argument meaning:

* ``test(model, test_dataset, indices=None, ...)``  -- ``learn.test`` (src/learn.py:409-459)
* ``compute_eer / compute_error_rates / compute_mindcf / get_test_metrics``  -- src/utils.py:294-403

The reference forwards BOTH utterances of every ordered pair through the model one at a time
(2 N^2 single-utterance forwards for N test utterances, src/learn.py:436-439 over
``get_sample_pairs``' ``itertools.product``, src/datasets.py:165-183) and then runs Python loops and
sklearn / scipy over the N^2 scores.  Here every utterance is embedded ONCE (eval mode: running
BatchNorm statistics, and the SE mean / attentive pooling are per utterance, so batching utterances
of EQUAL frame count changes nothing; different lengths are never padded together because the
reference does not mask), the N x N cosine scores and same-speaker labels come from one kernel
(``tn_cosine_scores``) and the sort / prefix counts / minimum detection cost / ROC crossing from
``tn_det_metrics``.  No CPU fallback: everything below raises on CPU tensors.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from ._lib import LIB, call, ptr, require_cuda

Tensor = torch.Tensor
COSINE_EPS = 1e-8          # F.cosine_similarity default (src/learn.py:438)


# ----------------------------------------------------------------------------
# scoring
# ----------------------------------------------------------------------------
def cosine_scores(embeddings: Tensor, speakers: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """``scores[i, j] = F.cosine_similarity(e_i, e_j)`` and ``labels[i, j] = speakers[i] == speakers[j]`` (uint8).

    Row-major flattening gives the trial list of ``get_sample_pairs`` (``itertools.product(indices, repeat=2)``)."""
    require_cuda(embeddings, speakers)
    if embeddings.dim() != 2:
        raise ValueError("embeddings must be [N, E]")
    e = embeddings.detach().to(torch.float32).contiguous()
    n, d = e.shape
    scores = torch.empty((n, n), device=e.device, dtype=torch.float32)
    labels = spk = None
    if speakers is not None:
        spk = speakers.detach().to(torch.int64).contiguous()
        if spk.shape != (n,):
            raise ValueError("speakers must be [N]")
        labels = torch.empty((n, n), device=e.device, dtype=torch.uint8)
    call("tn_cosine_scores", ptr(e), ptr(spk), ptr(scores), ptr(labels), n, d, COSINE_EPS)
    return scores, labels


@torch.no_grad()
def embed_utterances(model: torch.nn.Module, spectrograms: Sequence[Tensor], max_batch: int = 256) -> Tensor:
    """Eval-mode embeddings ``[N, E]`` of N spectrograms (``[1, n_mels, T_i]`` or ``[n_mels, T_i]``), each forwarded once.

    Utterances are bucketed by frame count and each bucket runs as a batch; the result equals the reference's
    one-utterance-at-a-time ``model(s)`` (src/learn.py:437) because nothing in eval mode mixes utterances."""
    model.eval()
    specs = [s if s.dim() == 3 else s.unsqueeze(0) for s in spectrograms]
    for s in specs:
        if s.dim() != 3 or s.shape[0] != 1:
            raise ValueError("each spectrogram must be [1, n_mels, T] or [n_mels, T]")
    buckets: Dict[int, List[int]] = {}
    for i, s in enumerate(specs):
        buckets.setdefault(int(s.shape[-1]), []).append(i)
    dev = next(model.parameters()).device
    out: Optional[Tensor] = None
    for _, idx in sorted(buckets.items()):
        for c0 in range(0, len(idx), max_batch):
            chunk = idx[c0:c0 + max_batch]
            x = torch.cat([specs[i] for i in chunk], dim=0).to(dev, torch.float32)
            emb = model(x)
            if out is None:
                out = torch.empty((len(specs), emb.shape[1]), device=dev, dtype=torch.float32)
            out[torch.as_tensor(chunk, device=dev)] = emb
    if out is None:
        raise ValueError("no utterances")
    return out


# ----------------------------------------------------------------------------
# detection metrics
# ----------------------------------------------------------------------------
def _device_trials(scores, labels) -> Tuple[Tensor, Tensor]:
    dev = None
    for t in (scores, labels):
        if isinstance(t, Tensor) and t.is_cuda:
            dev = t.device
    if dev is None:
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    if dev is None:
        require_cuda(torch.empty(0))        # raises: there is no CPU path
    s = torch.as_tensor(scores).detach().to(dev, torch.float32).reshape(-1).contiguous()
    lab = torch.as_tensor(labels).detach().to(dev)
    lab = (lab != 0).to(torch.uint8).reshape(-1).contiguous()
    if s.numel() != lab.numel() or s.numel() == 0:
        raise ValueError("scores and labels must be non-empty and of equal length")
    require_cuda(s, lab)
    return s, lab


class DetResult:
    """Raw outputs of ``tn_det_metrics`` for one trial list (device tensors; ``.item()`` them to sync)."""

    def __init__(self, out8: Tensor, fnrs: Optional[Tensor], fprs: Optional[Tensor], order: Optional[Tensor],
                 p_target: float, c_fa: float, c_miss: float, eps: float):
        self.out8, self.fnrs, self.fprs, self.order = out8, fnrs, fprs, order
        self.p_target, self.c_fa, self.c_miss, self.eps = p_target, c_fa, c_miss, eps

    @property
    def eer(self) -> float:
        return float(self.out8[1].item())

    @property
    def mindcf(self) -> float:
        # min_dcf = min_c_det / (c_def + eps), c_def = min(c_miss * p_target, c_fa * (1 - p_target))   (src/utils.py:366-370)
        c_def = min(self.c_miss * self.p_target, self.c_fa * (1 - self.p_target))
        return float(self.out8[0].item()) / (c_def + self.eps)


def det_metrics(scores, labels, p_target: float = 1e-2, c_fa: float = 1, c_miss: float = 1, eps: float = 1e-6,
                want_rates: bool = False, want_order: bool = False) -> DetResult:
    """One pass of ``tn_det_metrics`` over a trial list (scores are taken as fp32, labels as 0/1)."""
    s, lab = _device_trials(scores, labels)
    n = s.numel()
    nbytes = ctypes.c_longlong(0)
    LIB.load()
    LIB.call("tn_det_workspace_bytes", n, ctypes.addressof(nbytes))
    ws = torch.empty(nbytes.value, device=s.device, dtype=torch.uint8)
    out8 = torch.empty(8, device=s.device, dtype=torch.float64)
    fnrs = torch.empty(n, device=s.device, dtype=torch.float64) if want_rates else None
    fprs = torch.empty(n, device=s.device, dtype=torch.float64) if want_rates else None
    keys = torch.empty(n, device=s.device, dtype=torch.int64) if want_order else None
    call("tn_det_metrics", ptr(s), ptr(lab), n, float(p_target), float(c_fa), float(c_miss), float(eps), ptr(ws),
         nbytes.value, ptr(out8), ptr(fnrs), ptr(fprs), ptr(keys))
    order = (keys & 0xFFFFFFFF) if keys is not None else None
    return DetResult(out8, fnrs, fprs, order, float(p_target), float(c_fa), float(c_miss), float(eps))


def compute_error_rates(scores, labels, eps: float = 1e-6) -> Tuple[List[float], List[float]]:
    """``utils.compute_error_rates`` (src/utils.py:303-350): false-negative / false-positive rates at every
    threshold of the ascending (stable) score order, as Python lists like the reference."""
    r = det_metrics(scores, labels, eps=eps, want_rates=True)
    return r.fnrs.tolist(), r.fprs.tolist()


def compute_mindcf(scores, labels, p_target: float = 1e-2, c_fa: float = 1, c_miss: float = 1, eps: float = 1e-6) -> float:
    """``utils.compute_mindcf`` (src/utils.py:353-372).  Like the reference, the rates inside use
    ``compute_error_rates``' default eps (1e-6) and ``eps`` only guards the final normalisation."""
    r = det_metrics(scores, labels, p_target=p_target, c_fa=c_fa, c_miss=c_miss, eps=1e-6)
    r.eps = float(eps)
    return r.mindcf


def compute_eer(scores, labels) -> float:
    """``utils.compute_eer`` (src/utils.py:294-300): abscissa where the ROC polyline meets ``tpr = 1 - fpr``."""
    return det_metrics(scores, labels).eer


def get_test_metrics(scores, labels, mindcf_p_target: float = 1e-2, mindcf_c_fa: float = 1, mindcf_c_miss: float = 1,
                     prefix: Optional[str] = None) -> Dict[str, float]:
    """``utils.get_test_metrics`` (src/utils.py:386-403): both metrics from ONE sort of the trials."""
    r = det_metrics(scores, labels, p_target=mindcf_p_target, c_fa=mindcf_c_fa, c_miss=mindcf_c_miss)
    metrics = {"eer": r.eer, "mindcf": r.mindcf}
    if prefix is not None:
        metrics = {f"{prefix}/{k}": v for k, v in metrics.items()}
    return metrics


# ----------------------------------------------------------------------------
# learn.test
# ----------------------------------------------------------------------------
def _dataset_items(test_dataset, indices: Optional[Iterable[int]]):
    if isinstance(test_dataset, torch.utils.data.Subset):          # src/learn.py:429-434
        test_dataset, indices = test_dataset.dataset, test_dataset.indices
    indices = list(indices) if indices else list(range(len(test_dataset)))   # `indices or range(len)` (src/datasets.py:171)
    return [test_dataset[i] for i in indices]


@torch.no_grad()
def test(model, test_dataset, indices=None, wandb_run=None, log_console=True, mindcf_p_target=0.01, mindcf_c_fa=1,
         mindcf_c_miss=1, device="cuda"):
    """``learn.test`` (src/learn.py:409-459): EER and minDCF over every ordered pair of test utterances.

    ``test_dataset[i]`` must return the reference's example dict (``"spectrogram"`` ``[1, n_mels, T]`` and
    ``"speaker"``).  Returns ``{"test/eer": ..., "test/mindcf": ...}``."""
    items = _dataset_items(test_dataset, indices)
    emb = embed_utterances(model, [it["spectrogram"] for it in items])
    ids: Dict[object, int] = {}
    spk = torch.tensor([ids.setdefault(_speaker_key(it["speaker"]), len(ids)) for it in items], dtype=torch.int64,
                       device=emb.device)
    scores, labels = cosine_scores(emb, spk)
    metrics = get_test_metrics(scores, labels, mindcf_p_target=mindcf_p_target, mindcf_c_fa=mindcf_c_fa,
                               mindcf_c_miss=mindcf_c_miss, prefix="test")
    if log_console:
        print("  ".join(f"{k}: {v:.6f}" for k, v in metrics.items()))
    if wandb_run is not None:                                       # src/learn.py:456-457
        import json
        wandb_run.notes = json.dumps(metrics, indent=2).encode("utf-8")
    return metrics


def _speaker_key(s):
    return s.item() if isinstance(s, Tensor) and s.numel() == 1 else s
