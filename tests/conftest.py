import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "stable-dreamfusion_b200")
for p in (ROOT, PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # a GPU test on a box without a GPU is an error in the harness, not a skip-worthy event,
    # but collecting them on the CPU box must not fail: they are deselected by -m "not gpu".
    pass


@pytest.fixture(scope="session")
def device():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")
