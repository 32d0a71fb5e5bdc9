"""Real multi-GPU run (NCCL, one process per GPU) of the sharded table, checked against the C oracle.  Needs two
GPUs on the box; skipped otherwise (the one-GPU loopback tests in test_gpu_distributed.py cover the same logic)."""
import os
import socket
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_table_over_nccl():
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "MULTI_GPU_CHECK PASS" in res.stdout
