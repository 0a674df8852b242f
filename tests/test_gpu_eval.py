"""GPU parity of the evaluation consumer (titanet_b200.evaluation -> tn_cosine_scores / tn_det_metrics, through the C ABI)
against the reference-generated goldens (tests/golden/eval_metrics.npz) and the numpy oracle (oracle/eval_oracle.py).

Bars: sort order, labels and counts bit-exact; fnrs / fprs / minDCF bit-exact fp64 (same operation order as the
reference's Python loops); EER within 1e-9 of scipy's brentq (xtol 2e-12); cosine scores within 1e-6 of fp32 torch."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import eval_oracle as E  # noqa: E402  (checker only)
from eval_cases import MODEL_SPEAKERS, MODEL_SPEC, model_utterances, trial_cases  # noqa: E402


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "eval_metrics.npz"))


@pytest.mark.parametrize("name", list(trial_cases()))
def test_det_metrics_match_reference(gold, name):
    from titanet_b200 import evaluation as ev
    scores, labels = trial_cases()[name]
    r = ev.det_metrics(scores, labels, want_rates=True, want_order=True)
    assert np.array_equal(r.order.cpu().numpy(), np.argsort(scores.astype(np.float64), kind="stable"))
    assert int(r.out8[2].item()) == int(labels.sum()) and int(r.out8[3].item()) == int(len(labels) - labels.sum())
    assert abs(r.eer - float(gold[f"{name}:eer"])) < 1e-9
    assert r.mindcf == pytest.approx(float(gold[f"{name}:mindcf"]), rel=1e-12, abs=0)
    fnrs, fprs, _ = E.compute_error_rates(scores, labels)
    assert np.array_equal(r.fnrs.cpu().numpy(), fnrs)
    assert np.array_equal(r.fprs.cpu().numpy(), fprs)
    if f"{name}:fnrs" in gold:
        assert np.array_equal(r.fnrs.cpu().numpy(), gold[f"{name}:fnrs"])
        assert np.array_equal(r.fprs.cpu().numpy(), gold[f"{name}:fprs"])
    # the reference-named entry points
    assert ev.compute_mindcf(scores, labels, p_target=0.05, c_fa=2, c_miss=3) == pytest.approx(
        float(gold[f"{name}:mindcf_p05"]), rel=1e-12, abs=0)
    m = ev.get_test_metrics(torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda(), prefix="test")
    assert set(m) == {"test/eer", "test/mindcf"}
    assert abs(m["test/eer"] - float(gold[f"{name}:eer"])) < 1e-9
    assert abs(ev.compute_eer(list(scores), list(labels)) - float(gold[f"{name}:eer"])) < 1e-9


def test_det_metrics_million_trials():
    """Full test-split size (1 000 utterances -> 10^6 ordered pairs): nine global bitonic stages, 512 scan chunks."""
    from titanet_b200 import evaluation as ev
    rng = np.random.RandomState(5)
    n = (1 << 20) + 37
    labels = (rng.rand(n) < 0.01).astype(np.int64)
    scores = np.round(np.clip(rng.randn(n) * 0.2 + 0.5 * labels, -1, 1), 4).astype(np.float32)    # many ties
    r = ev.det_metrics(scores, labels, want_rates=True, want_order=True)
    fnrs, fprs, order = E.compute_error_rates(scores, labels)
    assert np.array_equal(r.order.cpu().numpy(), order)
    assert np.array_equal(r.fnrs.cpu().numpy(), fnrs) and np.array_equal(r.fprs.cpu().numpy(), fprs)
    assert abs(r.eer - E.compute_eer(scores, labels)) < 1e-12
    assert r.mindcf == pytest.approx(E.compute_mindcf(scores, labels), rel=1e-12, abs=0)


def test_det_metrics_single_class_and_single_trial():
    from titanet_b200 import evaluation as ev
    r = ev.det_metrics(np.asarray([0.1, 0.7, 0.3], np.float32), np.asarray([1, 1, 1]))
    assert np.isnan(r.eer)                     # the reference's brentq has no bracket here and raises
    assert r.mindcf == pytest.approx(E.compute_mindcf([0.1, 0.7, 0.3], [1, 1, 1]), rel=1e-12)
    r = ev.det_metrics([0.5], [0])
    assert np.isnan(r.eer) and int(r.out8[3].item()) == 1
    with pytest.raises(ValueError):
        ev.det_metrics([], [])


@pytest.mark.parametrize("n,d", [(7, 48), (70, 192), (33, 100)])
def test_cosine_scores(gold, n, d):
    from titanet_b200 import evaluation as ev
    g = torch.Generator().manual_seed(n)
    e = torch.randn(n, d, generator=g)
    e[1] = 0.0                                  # zero vector: the eps clamp, score 0
    spk = torch.randint(0, 4, (n,), generator=g)
    s, lab = ev.cosine_scores(e.cuda(), spk.cuda())
    ref = torch.nn.functional.cosine_similarity(e.double()[:, None, :], e.double()[None, :, :], dim=2)
    assert (s.cpu().double() - ref).abs().max() < 1e-6
    assert torch.equal(lab.cpu().bool(), spk[:, None] == spk[None, :])
    so, lo = E.sample_pair_trials(e.numpy(), spk.numpy())
    assert np.abs(s.cpu().numpy().reshape(-1) - so).max() < 1e-6 and np.array_equal(lab.cpu().numpy().reshape(-1), lo)
    s2, none = ev.cosine_scores(e.cuda())
    assert none is None and torch.equal(s2, s)


def test_learn_test_matches_reference_loop(gold):
    """learn.test (src/learn.py:409-459) over a 7-utterance split of 4 different lengths: every utterance embedded once in
    per-length batches vs the reference's 2 x 49 single-utterance forwards."""
    from test_gpu_model import build_model
    from titanet_b200 import evaluation as ev
    model = build_model(MODEL_SPEC)
    specs = model_utterances()
    data = [{"spectrogram": s, "speaker": f"spk{k}"} for s, k in zip(specs, MODEL_SPEAKERS)]
    emb = ev.embed_utterances(model, specs)
    assert not model.training
    assert (emb.cpu() - torch.from_numpy(gold["model:emb"])).abs().max() < 1e-3
    one_by_one = torch.cat([model(s.cuda()) for s in specs])
    assert (emb - one_by_one).abs().max() < 1e-5          # batching by length changes nothing
    spk = torch.tensor(MODEL_SPEAKERS).cuda()
    scores, labels = ev.cosine_scores(emb, spk)
    assert np.array_equal(labels.cpu().numpy().reshape(-1), gold["model:labels"])
    assert np.abs(scores.cpu().numpy().reshape(-1) - gold["model:scores"]).max() < 1e-3
    m = ev.test(model, data, log_console=False)
    s_np, l_np = scores.cpu().numpy().reshape(-1), labels.cpu().numpy().reshape(-1)
    assert abs(m["test/eer"] - E.compute_eer(s_np, l_np)) < 1e-9
    assert m["test/mindcf"] == pytest.approx(E.compute_mindcf(s_np, l_np), rel=1e-12)
    assert abs(m["test/eer"] - float(gold["model:eer"])) < 0.05 and abs(m["test/mindcf"] - float(gold["model:mindcf"])) < 0.05
    sub = torch.utils.data.Subset(data, [0, 1, 2, 4])
    ms = ev.test(model, sub, log_console=False)
    e4 = emb[[0, 1, 2, 4]]
    s4, l4 = ev.cosine_scores(e4, spk[[0, 1, 2, 4]])
    assert abs(ms["test/eer"] - E.compute_eer(s4.cpu().numpy().reshape(-1), l4.cpu().numpy().reshape(-1))) < 1e-9


def test_evaluation_rejects_cpu_tensors():
    from titanet_b200 import TitanetLibraryError, evaluation as ev
    with pytest.raises(TitanetLibraryError):
        ev.cosine_scores(torch.randn(4, 8))
