"""Multi-GPU value parity: 2 real GPUs, NCCL, one process per GPU.  Skipped on a one-GPU box; run with
``gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp2.py -m gpu`` (log under profiles/)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_step_matches_per_shard_oracle():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "dp2_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0 and len(lines) == 2 and all(l["ok"] for l in lines), (res.stdout[-2000:], res.stderr[-2000:])
