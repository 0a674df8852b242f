"""bench.py contract on a machine without a GPU: the reference arm prints ONE JSON line with the keys the driver reads; our arm
refuses to run (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    res = run("--impl", "reference", "--blocks", "1", "--cpu-batch", "2", "--steps", "1", "--warmup", "1", "--seconds", "1")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "utterances/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "utterances/sec (TitaNet-S fwd+bwd, 1s@16kHz)" and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["config"]["workload"].startswith("TitaNet-S/1 fwd+bwd, CE loss")
    cb = d["cpu_baseline"]
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "models.py"))      # installed by __graft_entry__.build()
    assert cb["kind"] == ("reference" if have_ref else "port")
    assert cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "utterances" in cb["sample"]
    assert d["config"]["batch_per_gpu"] == 2 and d["reference_default_threads"]["cores"] == 2
    assert d["e2e"] == {"value": d["value"], "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=300, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_our_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    res = run("--steps", "1", "--warmup", "1")
    assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)
