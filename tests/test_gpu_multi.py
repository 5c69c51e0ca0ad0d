"""Multi-GPU parity (skipped with fewer than two GPUs): a torchrun-spawned 2-rank NCCL job compares the
exchanged factors of ``FactorSweep`` with a single-GPU evaluation of the concatenated batch."""

import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_factor_exchange_matches_single_gpu():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29531", str(ROOT / "tests" / "multi" / "nccl_factor_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=str(ROOT))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "nccl factor parity ok" in res.stdout
