"""NCCL test (needs >= 2 GPUs; skipped otherwise): TFIDF(distributed=True) under torchrun returns, on every rank,
frames identical to the single-GPU matcher (tools/dist_check.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_match_equals_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tools", "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "DIST_CHECK PASS" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
