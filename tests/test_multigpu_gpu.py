"""z-slab sharding over NCCL (2 ranks): SpMV, solve, filter and MMA through the C-ABI give the same results
as the single-GPU run.  Skipped when fewer than 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_parity():
    import torch
    if torch.cuda.device_count() < 2:
        print("SKIPPED: 1 GPU visible -- the 2-rank parity test needs 2 (it runs wherever >= 2 GPUs are visible)")
        pytest.skip("SKIPPED: 1 GPU visible, the 2-rank parity test needs 2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MP_PARITY_OK" in r.stdout
