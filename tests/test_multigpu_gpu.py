"""Launches tests/multigpu_check.py (2 ranks x B == 1 rank x 2B on the product path, bit-identical replicas after
graph-replayed steps) under torchrun when the box has at least two GPUs; skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs on one box")
def test_two_ranks_of_the_cuda_path_equal_one_rank_with_the_whole_batch():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29511", os.path.join(ROOT, "tests", "multigpu_check.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MULTIGPU_CHECK" in p.stdout and '"replicas_identical": true' in p.stdout
