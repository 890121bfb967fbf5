"""-m gpu, needs >= 2 GPUs (skipped otherwise): the partitioned path over NCCL — ghost-layer assembly,
halo exchange, allreduce inner products, per-rank ILU(0) — against the 1-rank CPU oracle."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_parity():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run: gpurun --gpus 2 -- python -m pytest tests -m gpu -k two_rank)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "mgpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    fails = [l for l in out.stdout.splitlines() if "FAIL" in l]
    assert "MGPU_CHECK PASS" in out.stdout, "\n".join(fails) + "\n--- stdout tail ---\n" + out.stdout[-2500:] + "\n--- stderr tail ---\n" + out.stderr[-2500:]
