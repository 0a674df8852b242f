"""CPU oracle for the evaluation consumer of the TitaNet path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/`` (and ``oracle/make_golden_eval.py``) import this module; ``titanet_b200.evaluation`` never does.

A numpy restatement of (paths relative to the reference checkout, Wadaboa/titanet @ 7b77053):

* the trial list of ``SpeakerDataset.get_sample_pairs`` (src/datasets.py:165-183: ``itertools.product(indices, repeat=2)``,
  label = same speaker) scored as ``learn.test`` does (src/learn.py:436-439: ``F.cosine_similarity(e1, e2)``);
* ``utils.compute_error_rates`` (src/utils.py:303-350), ``utils.compute_mindcf`` (353-372), ``utils.compute_eer`` (294-300).

Third-party arithmetic restated (absent from the reference tree, pinned by ``init/requirements.txt``):
``sklearn.metrics.roc_curve`` (scikit-learn 1.1.3) -- ROC points at the distinct scores in descending order, origin
prepended, collinear points optionally dropped (which does not change the polyline) -- and
``scipy.optimize.brentq`` over ``scipy.interpolate.interp1d(fpr, tpr)`` (scipy 1.9.3), i.e. the abscissa where that
polyline meets ``tpr = 1 - fpr``; here it is found in closed form on the crossing segment.

Parity pinning: ``oracle/make_golden_eval.py`` executes the reference's OWN function bodies (extracted from
``/root/reference/src/utils.py`` with ``ast``; the module itself does not import here -- matplotlib / umap / IPython are
missing) with the installed scikit-learn / scipy and commits the results as ``tests/golden/eval_metrics.npz``;
``tests/test_eval_oracle.py`` holds this file to those numbers.
"""
from __future__ import annotations

import itertools

import numpy as np


def cosine_similarity(e1: np.ndarray, e2: np.ndarray, eps: float = 1e-8) -> np.float32:
    """``F.cosine_similarity`` on two vectors, fp32 (each vector divided by ``max(norm, eps)``, then the dot product)."""
    e1 = np.asarray(e1, np.float32)
    e2 = np.asarray(e2, np.float32)
    n1 = max(np.float32(np.sqrt(np.sum(e1 * e1, dtype=np.float32))), np.float32(eps))
    n2 = max(np.float32(np.sqrt(np.sum(e2 * e2, dtype=np.float32))), np.float32(eps))
    return np.float32(np.sum((e1 / n1) * (e2 / n2), dtype=np.float32))


def sample_pair_trials(embeddings: np.ndarray, speakers):
    """scores, labels of every ORDERED pair (i1, i2) in ``itertools.product`` order (src/datasets.py:171-182, src/learn.py:436-439)."""
    n = len(embeddings)
    scores, labels = [], []
    for i1, i2 in itertools.product(range(n), repeat=2):
        scores.append(cosine_similarity(embeddings[i1], embeddings[i2]))
        labels.append(int(speakers[i1] == speakers[i2]))
    return np.asarray(scores, np.float32), np.asarray(labels, np.int64)


def compute_error_rates(scores, labels, eps: float = 1e-6):
    """src/utils.py:303-350.  Stable ascending sort by score; ``fnrs[i]`` = targets at or below threshold i / (targets + eps),
    ``fprs[i]`` = 1 - non-targets at or below threshold i / (non-targets + eps)."""
    scores = np.asarray(scores, np.float64)
    labels = np.asarray(labels).astype(np.int64)
    order = np.argsort(scores, kind="stable")
    lab = labels[order]
    fn = np.cumsum(lab)
    fp = np.cumsum(1 - lab)
    fnrs = fn / (float(lab.sum()) + eps)
    fprs = 1 - fp / (float(len(lab) - lab.sum()) + eps)
    return fnrs, fprs, order


def compute_mindcf(scores, labels, p_target=1e-2, c_fa=1, c_miss=1, eps=1e-6):
    """src/utils.py:353-372."""
    fnrs, fprs, _ = compute_error_rates(scores, labels)
    c_det = c_miss * fnrs * p_target + c_fa * fprs * (1 - p_target)
    c_def = min(c_miss * p_target, c_fa * (1 - p_target))
    return float(c_det.min()) / (c_def + eps)


def roc_points(scores, labels):
    """``sklearn.metrics.roc_curve(labels, scores, drop_intermediate=False)``: (fpr, tpr) with the origin first."""
    scores = np.asarray(scores, np.float64)
    labels = np.asarray(labels).astype(np.int64)
    order = np.argsort(-scores, kind="stable")
    s, lab = scores[order], labels[order]
    last_of_group = np.r_[np.nonzero(np.diff(s))[0], len(s) - 1]
    tps = np.cumsum(lab)[last_of_group]
    fps = (1 + last_of_group) - tps
    tps, fps = np.r_[0, tps], np.r_[0, fps]
    return fps / fps[-1], tps / tps[-1]


def compute_eer(scores, labels) -> float:
    """src/utils.py:294-300: root of ``1 - x - interp1d(fpr, tpr)(x)`` on [0, 1]."""
    fpr, tpr = roc_points(scores, labels)
    h = fpr + tpr                                   # non-decreasing along the curve, 0 at the origin, 2 at the end
    k = int(np.argmax(h >= 1.0))                    # first point on / above the anti-diagonal
    x0, y0, x1, y1 = fpr[k - 1], tpr[k - 1], fpr[k], tpr[k]
    t = (1.0 - x0 - y0) / ((x1 - x0) + (y1 - y0))
    return float(x0 + t * (x1 - x0))
