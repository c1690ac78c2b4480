import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (oracle/oracle.c) -- the checker, never the product."""
    import oracle
    return oracle.Oracle()


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_cpu.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def dr():
    """The product: ctypes binding of libdrjit_core_b200.so (CUDA only)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import drjit_core_b200 as dr
    dr.jit_init()
    return dr
