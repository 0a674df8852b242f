"""Generate tests/golden/eval_metrics.npz by running the UNMODIFIED reference here (build container only).

``python oracle/make_golden_eval.py``.  ``/root/reference/src/utils.py`` does not import in this image (matplotlib,
umap, IPython are missing), so the four metric functions are taken out of its source with ``ast`` and executed as
they are, against the installed scikit-learn / scipy (the reference pins 1.1.3 / 1.9.3).  The ``learn.test`` case
runs the reference's own ``models.TitaNet`` in eval mode one utterance at a time over
``itertools.product(indices, repeat=2)`` exactly as src/learn.py:436-439 / src/datasets.py:171-182 do
(``learn`` / ``datasets`` do not import either).  TEST INFRASTRUCTURE.
"""
import ast
import itertools
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get("TITANET_REFERENCE", "/root/reference/src")
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import titanet_oracle as O  # noqa: E402
from eval_cases import MODEL_SPEAKERS, MODEL_SPEC, model_utterances, trial_cases  # noqa: E402
from make_golden import build_ref  # noqa: E402  (imports the reference's modules / models / losses)


def reference_metric_functions():
    from scipy.interpolate import interp1d
    from scipy.optimize import brentq
    from sklearn.metrics import roc_curve
    src = open(os.path.join(REF, "utils.py")).read()
    wanted = {"compute_eer", "compute_error_rates", "compute_mindcf", "get_test_metrics"}
    tree = ast.parse(src)
    tree.body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    ns = dict(roc_curve=roc_curve, brentq=brentq, interp1d=interp1d)
    exec(compile(tree, os.path.join(REF, "utils.py"), "exec"), ns)
    assert wanted <= set(ns)
    return ns


def main():
    ref = reference_metric_functions()
    out = {}
    for name, (scores, labels) in trial_cases().items():
        s, lab = [float(v) for v in scores], [int(v) for v in labels]       # learn.test builds Python lists (.item())
        m = ref["get_test_metrics"](s, lab)
        out[f"{name}:eer"], out[f"{name}:mindcf"] = m["eer"], m["mindcf"]
        out[f"{name}:mindcf_p05"] = ref["compute_mindcf"](s, lab, p_target=0.05, c_fa=2, c_miss=3)
        if len(s) <= 3000:
            fnrs, fprs = ref["compute_error_rates"](s, lab)
            out[f"{name}:fnrs"], out[f"{name}:fprs"] = np.asarray(fnrs), np.asarray(fprs)
        print(name, m)

    # learn.test over a tiny split
    model = build_ref(MODEL_SPEC, None, 0).eval()
    specs = model_utterances()
    scores, labels = [], []
    with torch.no_grad():
        for i1, i2 in itertools.product(range(len(specs)), repeat=2):
            e1, e2 = model(specs[i1]), model(specs[i2])
            scores += [F.cosine_similarity(e1, e2).item()]
            labels += [int(MODEL_SPEAKERS[i1] == MODEL_SPEAKERS[i2])]
        emb = torch.cat([model(s) for s in specs], dim=0)
    m = ref["get_test_metrics"](scores, labels, prefix="test")
    out["model:emb"], out["model:scores"], out["model:labels"] = emb.numpy(), np.asarray(scores, np.float32), np.asarray(labels)
    out["model:eer"], out["model:mindcf"] = m["test/eer"], m["test/mindcf"]
    print("model", m)
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "eval_metrics.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
