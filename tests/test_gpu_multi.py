"""Multi-GPU paths on a box with >= 2 GPUs: the sharded NTT over NCCL (skipped on single-GPU boxes)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dist_ntt_matches_oracle(world):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "dist_ntt_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "dist_ntt OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
