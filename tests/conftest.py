import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One library context on cuda:0.  Fails (does not skip) when the CUDA path is unusable:
    a silent fallback would void every parity claim."""
    import shielded_pool_pinocchio_solana_b200 as g16
    c = g16.Context(0)
    yield c
    c.close()
