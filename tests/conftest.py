import os
import sys

import pytest

# the peer-composite tests run several contexts (two streams each) with spin-waiting kernels on ONE device: give
# every stream its own hardware queue so that a waiting kernel cannot sit in front of the kernel it waits for
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU oracle (test infrastructure).  Builds oracle/libspim_oracle.so on first use."""
    from oracle import build, oracle
    build.build_oracle()
    build.build_ref()
    return oracle


@pytest.fixture(scope="session")
def have_ref(oracle_mod):
    return oracle_mod.available("reference")
