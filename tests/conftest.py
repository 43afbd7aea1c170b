"""pytest configuration: the ``gpu`` marker and shared fixtures."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def unhex(v):
    """Inverse of oracle.gen_golden.hx."""
    if isinstance(v, list):
        return [unhex(e) for e in v]
    if isinstance(v, str):
        return float.fromhex(v)
    return v


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "reference_kat.json")) as f:
        return json.load(f)


def ellipse_path(n, a, b):
    """The ellipse recipe of the golden vectors (oracle/gen_golden.py)."""
    import numpy as np
    th = np.linspace(0, 2 * np.pi, n, endpoint=False)
    return a * np.cos(th), b * np.sin(th)


GOLDEN_PATHS = {"ellipse720": (720, 40.0, 20.0), "ellipse400": (400, 30.0, 15.0)}


class Obj:
    """Duck-typed state object, as the reference's controllers accept (any object with .x .y .yaw .v)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


@pytest.fixture(scope="session")
def cuda_device():
    """cuda:0, or fail loudly: the ``gpu`` tests must never pass on a machine without the CUDA path."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked tests need a CUDA device (there is no CPU fallback)"
    return torch.device("cuda:0")
