"""The C-ABI library loads and exports every symbol include/g16b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "g16b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(g16_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    import shielded_pool_pinocchio_solana_b200 as g16
    names = declared_functions()
    assert len(names) >= 25
    lib = ctypes.CDLL(g16.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in g16b200.h but not exported: %s" % missing
    unbound = [n for n in names if n not in g16.PROTOTYPES]
    assert not unbound, "no ctypes prototype for: %s" % unbound


def test_no_gpu_means_loud_failure_not_fallback():
    """On a box without a GPU every compute entry point must fail with G16_E_CUDA (4)."""
    import torch
    import shielded_pool_pinocchio_solana_b200 as g16
    if torch.cuda.is_available():
        return
    try:
        g16.Context(0)
    except g16.G16Error as e:
        assert e.code == 4 and "no CPU fallback" in str(e)
    else:
        raise AssertionError("Context(0) succeeded without a GPU")
