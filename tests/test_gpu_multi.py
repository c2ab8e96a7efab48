"""Slab-partitioned multi-GPU parity: runs tests/mgpu_check.py under torchrun on 2 GPUs when the box has them."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, port):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", script)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(p.stdout[-4000:])
    sys.stderr.write(p.stderr[-4000:])
    assert p.returncode == 0


def test_two_gpu_operators_match_oracle():
    """reductions, halo'd stencils and the transposed Green operator of every mode / scheme against the CPU oracle"""
    _torchrun("mgpu_ops.py", 29533)


def test_two_gpu_slab_partition_matches_single_gpu():
    """same iteration count / residual history / mean stress / strain field as the single-GPU solve of the same problem"""
    _torchrun("mgpu_check.py", 29531)
