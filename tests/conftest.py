import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("python-bls_b200", "oracle", ""):
    p = os.path.join(ROOT, sub)
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def unhex_elems(h, width=48):
    """hex of k*48 bytes -> tuple of k ints"""
    raw = bytes.fromhex(h)
    return tuple(int.from_bytes(raw[i:i + width], "big") for i in range(0, len(raw), width))
