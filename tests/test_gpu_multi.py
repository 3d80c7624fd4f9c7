"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the train-sharded matcher over peer memory and over
NCCL equals a single-device pass -- tests/mp_sharded_check.py under torchrun, one process per GPU.

Not run with several ranks on ONE GPU: kernels of different processes that wait on each other's flags are not guaranteed
to run at the same time there (B200_PROFILING.md: context-switch timeouts).  On a 1-GPU box the exchange is covered by
tests/test_gpu_matcher.py::test_p2p_fused_scatter_merge_single_process (ranks emulated by handles of one process, no kernel
waiting on a later launch), and every multi-GPU run of bench.py checks the sharded result itself (hamming_parity_ok)."""
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
