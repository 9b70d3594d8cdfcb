import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def _has_gpu():
    try:
        import viamd_b200 as vb
        return vb.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def ref_harness():
    """Path of the unmodified-reference harness (strict build) or skip. Prebuilt files travel to the GPU box."""
    p = os.path.join(ROOT, "oracle", "_ref", "ref_harness_strict")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/ref_harness_strict not built (needs /root/reference: make -C oracle ref)")
    return p
