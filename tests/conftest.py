import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def mrag():
    import mrag_b200
    return mrag_b200


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.clib()
    return o


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly (not skip) on a GPU box without the library; on a CPU-only box they are
    # deselected by `-m "not gpu"`, and skipped if someone runs the whole suite there.
    if has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
