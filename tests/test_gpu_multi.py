"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the train-sharded matcher over peer memory and over
NCCL equals a single-device pass -- tests/mp_sharded_check.py under torchrun, one process per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_matcher_multi_process():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = min(n, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mp_sharded_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MP_SHARDED_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
    print(out.stdout.strip().splitlines()[-1])
