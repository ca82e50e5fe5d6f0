"""GPU (>= 2 devices): sharded search over real NCCL / NVLink -- fused peer-memory exchange == all-gather path
== single-GPU answer.  Skipped on a single-GPU box; the gloo tests cover the host logic on CPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_search_two_gpus(cuda_device):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
