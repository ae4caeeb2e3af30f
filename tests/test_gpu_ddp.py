"""Data-parallel training step over NCCL (rows shard, gradients all-reduce): needs >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_ddp.py -m gpu`); skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["ddp", "overlap"])
def test_ddp_gradients_match_whole_batch(mode):
    """mode "ddp": the reference's setup, DDP's bucketed all-reduce after the native backward; mode "overlap":
    comm.enable_overlapped_grad_sync — NCCL all-reduce per gradient group from inside the backward."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29517" if mode == "ddp" else "29518",
           os.path.join(ROOT, "tests", "helpers", "ddp_train_worker.py"), mode]
    env = dict(os.environ, NCCL_DEBUG="WARN")
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DDP_TRAIN_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
