import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    """The CPU checkers are test infrastructure: (re)build the port always, oracle/_ref when the
    reference mount exists (otherwise the prebuilt files that travelled with the snapshot are used)."""
    from oracle import pyoracle
    pyoracle.build("all")
    pyoracle.build("dropin")  # the drop-in binaries of test_dropin / test_replay: never test a stale shim (no-op without /root/reference)
    yield
