"""Multi-GPU checks, run only on a box with at least two GPUs (gpurun --gpus 2 / 8): the block-decomposed similarity scan of
`cosine_pairs_sharded` over NCCL finds exactly the pairs of the single-GPU scan of the same data."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_similarity_equals_single_gpu_scan():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "check_sharded_similarity.py"), "--rows-per-gpu", "16384"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "equal: True" in out.stdout
