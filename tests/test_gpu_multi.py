"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the train-sharded matcher over peer memory and over
NCCL equals a single-device pass -- tests/mp_sharded_check.py under torchrun, one process per GPU."""
import os
import signal
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, env_extra, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mp_sharded_check.py")]
    env = dict(os.environ, **env_extra)
    # own process group, killed as a whole on a timeout: a rank left spinning on a device-side flag must not outlive the test
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, start_new_session=True)
    try:
        stdout, stderr = proc.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        stdout, stderr = proc.communicate()
        raise AssertionError("timed out after %d s\n%s\n%s" % (timeout, stdout[-2000:], stderr[-4000:]))
    assert proc.returncode == 0 and "MP_SHARDED_OK" in stdout, stdout[-2000:] + stderr[-4000:]
    print(stdout.strip().splitlines()[-1])


def test_sharded_matcher_multi_process():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    _run(min(n, 8), {}, 600)


def test_sharded_matcher_two_processes_one_gpu():
    """The same exchange between two PROCESSES that share GPU 0 (what a 1-GPU box can run): cudaIpc-mapped gather buffers,
    device-side flags, kernels of the two contexts time-sliced by the driver."""
    _run(2, {"ORBX_MP_SAME_DEVICE": "1"}, 240)
