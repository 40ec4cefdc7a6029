"""2-GPU data-parallel step over NCCL: needs two visible GPUs (`gpurun --gpus 2`), skipped otherwise.
The checks live in tests/mp_step_worker.py (one process per GPU, launched like the driver launches bench.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_rank_nccl_step_matches_oracle_and_ranks_agree():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29633", os.path.join(HERE, "mp_step_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if out.returncode != 0:
        print(out.stdout[-1500:])
        print(out.stderr[-4000:])
    assert out.returncode == 0, "worker failed (output printed above)"
    assert out.stdout.count(" ok: ") == 2, out.stdout
    print(out.stdout)
