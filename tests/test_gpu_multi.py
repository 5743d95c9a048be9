"""Sharded (multi-GPU) parity: runs tests/mgpu_check.py under torchrun on 2 GPUs when the box has
them.  On a 1-GPU box this is skipped; the host-side sharding logic is covered on CPU by
tests/test_sharding.py (gloo, world_size 2)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_gpu_run_matches_reference_golden():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "mgpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MGPU_OK" in res.stdout


def test_two_gpu_run_with_the_peer_memory_allreduce():
    """Same check with the CUDA-IPC / NVLink peer-memory all-reduce kernels instead of NCCL
    (opt-in, VBNMF_PEER_ALLREDUCE=1): the result must not change."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29733",
           os.path.join(ROOT, "tests", "mgpu_check.py")]
    env = dict(os.environ, VBNMF_PEER_ALLREDUCE="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MGPU_OK" in res.stdout
