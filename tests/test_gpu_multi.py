"""Multi-GPU tests (skipped with fewer than 2 GPUs): depth-slab decomposition of one volume over
NCCL, each rank checked against the unsplit computation."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
from conftest import ROOT  # noqa: E402


@pytest.mark.parametrize("world", [2, 4])
def test_depth_slab_forward_and_sampler_match_unsplit(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "slab_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
