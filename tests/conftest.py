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
