"""GPU (>= 2 devices): launches tools/multigpu_check.py under torchrun -- row-sharded FPS over NCCL and item-sharded
KNN must equal the single-GPU results bit for bit.  Skipped on single-GPU boxes."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
def test_sharded_paths_equal_single_gpu():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n = min(_ngpu(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTIGPU_CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
