"""Two-GPU sharded run against the CPU oracle (tests/mgpu_check.py under torchrun); skipped with fewer than
two visible CUDA devices.  The N>1 host logic is also covered on CPU by tests/test_distributed_cpu.py (gloo)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_gpu_sharded_run_matches_oracle():
    from conftest import _gpu_count
    if _gpu_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(HERE, "mgpu_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0 and "MGPU CHECK PASSED" in r.stdout
