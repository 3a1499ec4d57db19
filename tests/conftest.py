"""pytest wiring: registers the ``gpu`` marker and makes the product package importable.

The package directory name (``3d-pose-estimation-with-previleged-information_b200``) is not a
valid Python identifier, so ``__graft_entry__.load_package()`` registers it as ``b2pose``.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def b2pose():
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def dev():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch.device("cuda:0")


def rel_err(a, b):
    """max|a-b| / max|b|  -- the 'relative' of BASELINE.json's tolerances."""
    import torch
    if torch.is_tensor(a):
        a = a.detach().float().cpu().numpy()
    if torch.is_tensor(b):
        b = b.detach().float().cpu().numpy()
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
