"""2-GPU check of the apex DistributedDataParallel / SyncBatchNorm replacements over NCCL (tools/ddp_parity.py): after one
step on different per-rank batches both ranks hold identical parameters, and the SyncBN global-batch statistics make
the 2 x 8-image step equal (to bf16 noise) to a single-process step on the concatenated 16-image batch.
Skipped on boxes with fewer than two GPUs (the CPU suite covers the same host logic over gloo, world size 2)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ddp_syncbn_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tools", "ddp_parity.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK" in out.stdout, out.stdout[-2000:]
