"""Multi-GPU exchange of the step on real devices: world sizes 2 and 8 through
torch.distributed.run, one process per GPU.  Skipped on a single-GPU box (the host-side sharding logic is
covered on CPU by tests/test_sharding.py with gloo)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 8])
def test_peer_exchange_every_rank_checks_the_gathered_summaries(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-4000:] + "\n" + r.stderr[-8000:]
    assert r.stdout.count("MGPU_OK") == world, r.stdout[-2000:]
