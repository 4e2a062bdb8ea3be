import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


def pytest_sessionfinish(session, exitstatus):
    """A test may have started a single-rank process group in this process (the peer-exchange test): close it."""
    td = sys.modules.get("torch.distributed")
    try:
        if td is not None and td.is_available() and td.is_initialized():
            td.destroy_process_group()
    except Exception:
        pass
