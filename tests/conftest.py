import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "audio-flow-rs_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu() -> bool:
    try:
        import audioflow
        import ctypes
        n = ctypes.c_int(0)
        audioflow.load_library().af_device_count(ctypes.byref(n))
        return n.value > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def af():
    import audioflow
    audioflow.init()
    return audioflow


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.lib()
    return oracle
