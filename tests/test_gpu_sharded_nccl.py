"""Row-sharded search over NCCL on 2 GPUs == one index over all rows.  Needs two
visible GPUs; skipped on a one-GPU box (the gloo tests in test_host_logic.py cover the plumbing)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_search_over_nccl_matches_one_index():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "check_sharded_nccl.py"), "300000", "128"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SHARDED_NCCL_PARITY OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
