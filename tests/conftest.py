import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds")


@pytest.fixture(scope="session", autouse=True)
def _built_checkers():
    """Make sure the CPU checkers exist (the oracle always; oracle/_ref only
    where /root/reference is available or a prebuilt copy travelled along)."""
    import oracle
    if not oracle.have_oracle():
        oracle.build("oracle")
    if not oracle.have_ref() and os.path.isdir("/root/reference/src"):
        oracle.build("ref")
    yield


def gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
