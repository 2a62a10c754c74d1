"""Data-parallel training step over NCCL (BASELINE config 4's exchange step).  Needs >= 2 GPUs on the box; with one
GPU it is skipped (the host logic is covered by the world_size-2 gloo test in tests/test_host_cpu.py)."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_nccl_data_parallel_step(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = os.path.join(ROOT, "tests", "_nccl_dp_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29761", script], capture_output=True,
                       text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "NCCL_DP_OK" in r.stdout
