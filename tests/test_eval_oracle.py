"""CPU: the evaluation oracle (oracle/eval_oracle.py) against tests/golden/eval_metrics.npz, i.e. against the reference's
own ``utils.compute_eer / compute_error_rates / compute_mindcf`` executed by oracle/make_golden_eval.py, and against the
reference model's one-pair-at-a-time ``learn.test`` loop."""
import os

import numpy as np
import pytest

import eval_oracle as E
from eval_cases import MODEL_SPEAKERS, trial_cases


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "eval_metrics.npz"))


@pytest.mark.parametrize("name", list(trial_cases()))
def test_metrics_match_reference(gold, name):
    scores, labels = trial_cases()[name]
    assert abs(E.compute_eer(scores, labels) - float(gold[f"{name}:eer"])) < 1e-9          # brentq xtol 2e-12
    assert E.compute_mindcf(scores, labels) == pytest.approx(float(gold[f"{name}:mindcf"]), rel=1e-12, abs=0)
    assert E.compute_mindcf(scores, labels, p_target=0.05, c_fa=2, c_miss=3) == pytest.approx(
        float(gold[f"{name}:mindcf_p05"]), rel=1e-12, abs=0)
    if f"{name}:fnrs" in gold:
        fnrs, fprs, _ = E.compute_error_rates(scores, labels)
        assert np.array_equal(fnrs, gold[f"{name}:fnrs"])
        assert np.array_equal(fprs, gold[f"{name}:fprs"])


def test_trial_list_matches_reference_loop(gold):
    scores, labels = E.sample_pair_trials(gold["model:emb"], MODEL_SPEAKERS)
    assert np.array_equal(labels, gold["model:labels"])
    assert np.abs(scores - gold["model:scores"]).max() < 5e-7
    assert abs(E.compute_eer(gold["model:scores"], labels) - float(gold["model:eer"])) < 1e-9
    assert E.compute_mindcf(gold["model:scores"], labels) == pytest.approx(float(gold["model:mindcf"]), rel=1e-12)


def test_eer_closed_form_matches_sklearn_and_scipy_on_random_trials():
    """The closed-form crossing of oracle/eval_oracle.compute_eer against the third-party chain the reference calls
    (sklearn.metrics.roc_curve -> scipy interp1d -> brentq, src/utils.py:294-300) on 60 random trial lists with heavy ties,
    and compute_error_rates against a literal transcription-free check of its definition (counts at or below each threshold)."""
    sk = pytest.importorskip("sklearn.metrics")
    from scipy.interpolate import interp1d
    from scipy.optimize import brentq
    rng = np.random.RandomState(2024)
    for _ in range(60):
        n = int(rng.randint(8, 400))
        labels = (rng.rand(n) < rng.uniform(0.1, 0.9)).astype(np.int64)
        if labels.min() == labels.max():
            labels[0], labels[1] = 0, 1
        scores = (rng.randn(n) * 0.3 + rng.uniform(0.0, 0.6) * labels).astype(np.float32)
        if rng.rand() < 0.7:
            scores = np.round(scores, int(rng.randint(0, 3))).astype(np.float32)
        fpr, tpr, _ = sk.roc_curve(labels, scores)
        want = brentq(lambda x: 1.0 - x - interp1d(fpr, tpr)(x), 0.0, 1.0)
        assert abs(E.compute_eer(scores, labels) - want) < 1e-9
        fnrs, fprs, order = E.compute_error_rates(scores, labels)
        s_sorted, l_sorted = scores[order], labels[order]
        assert np.all(np.diff(s_sorted) >= 0)
        i = int(rng.randint(0, n))
        assert fnrs[i] == l_sorted[: i + 1].sum() / (labels.sum() + 1e-6)
        assert fprs[i] == 1 - (1 - l_sorted[: i + 1]).sum() / ((1 - labels).sum() + 1e-6)
