"""GPU, 2 ranks over NCCL: the C++ shared tracker (include/lmot_shared.h, host/shared_tracker.cpp) == ONE reference tracker fed the
concatenation (streams mode) / the frames in order (frames mode), and every rank ends each tick with the owner's track table
(tests/nccl_shared_worker.py).  Skips on a single-GPU box; the driver's multi-GPU tier and `gpurun --gpus 2` run it."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_share_one_track_table_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    env.setdefault("NCCL_DEBUG", "WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29741", os.path.join(ROOT, "tests", "nccl_shared_worker.py")], cwd=ROOT, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "NCCL_SHARED_OK" in r.stdout, r.stdout[-4000:]
