import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def _have_gpu() -> bool:
    try:
        from structure_from_motion_b200 import _native

        return _native.load_library().sfm_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    gpu = _have_gpu()
    from oracle import reference_shims

    ref = reference_shims.reference_available()
    for item in items:
        if "gpu" in item.keywords and not gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))


@pytest.fixture(scope="session")
def engine():
    from structure_from_motion_b200 import _native

    return _native.get_engine(0)
