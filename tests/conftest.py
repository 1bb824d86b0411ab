import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Lazy view over one npz fixture: g['case/key'] and g.case('case') -> dict."""

    def __init__(self, fname):
        self._z = np.load(os.path.join(GOLDEN, fname))

    def names(self):
        return sorted({k.split("/")[0] for k in self._z.files})

    def case(self, name):
        pre = name + "/"
        return {k[len(pre):]: self._z[k] for k in self._z.files if k.startswith(pre)}


@pytest.fixture(scope="session")
def core_cases():
    return Golden("core_cases.npz")


@pytest.fixture(scope="session")
def module_cases():
    return Golden("module_cases.npz")


@pytest.fixture(scope="session")
def setc_cases():
    return Golden("network_setC.npz")


def rel_err(a, b):
    """The parity metric of SURVEY.md s8(a): max|a-b| / max|b|."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
