import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))


def _ensure_built():
    """The shared objects are build artefacts (git-ignored): compile them on first use."""
    lib = os.path.join(ROOT, "shielded_pool_pinocchio_solana_b200", "libg16b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


def pytest_configure(config):
    _ensure_built()
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One library context on cuda:0.  Fails (does not skip) when the CUDA path is unusable:
    a silent fallback would void every parity claim."""
    import shielded_pool_pinocchio_solana_b200 as g16
    c = g16.Context(0)
    yield c
    c.close()
