"""Multi-GPU exchange of the step on real devices: world size 2 through
torch.distributed.run, one process per GPU.  Skipped on a single-GPU box (the host-side sharding logic is
covered on CPU by tests/test_sharding.py with gloo)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_exchange_world2_both_ranks_check_the_gathered_summaries():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-4000:] + "\n" + r.stderr[-8000:]
    assert r.stdout.count("MGPU_OK") == 2, r.stdout[-2000:]
