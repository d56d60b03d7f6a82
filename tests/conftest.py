import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def port():
    from oracle.port import Port
    return Port()


@pytest.fixture(scope="session")
def reflib():
    from oracle import refbind
    if not refbind.available(False):
        pytest.skip("oracle/_ref/libpoyref.so not built (needs /root/reference)")
    return refbind.RefLib(False)


@pytest.fixture(scope="session")
def ctx():
    import poy5_b200 as pb
    c = pb.Context(0)
    yield c
    c.close()
