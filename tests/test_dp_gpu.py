"""N > 1 on real GPUs (skipped with fewer than two): NCCL data parallelism == one process on the whole batch, replicas stay
bit-identical. The world-size-2 gloo tests in test_distributed_cpu.py cover the same host logic without a GPU."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_parity_and_identical_replicas():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", str(ROOT / "tools" / "dp_parity.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "dp_parity: PASS" in out.stdout
