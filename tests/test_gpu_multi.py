"""Two-GPU check of the peer-memory gradient exchange (skipped on single-GPU boxes)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_exchange_matches_nccl_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29700 + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "peer_exchange_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "peer exchange check: ok=True" in out.stdout
