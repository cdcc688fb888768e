import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def ref_available():
    """oracle/_ref (the reference's own TU on the OpenCV shim): built here, prebuilt on the GPU box."""
    from oracle import ref_lib
    try:
        ref_lib.build()
    except Exception:
        pass
    if not ref_lib.available():
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    return ref_lib
