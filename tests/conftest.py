import os, sys
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_devices():
    try:
        from gnina_b200 import capi
        return int(capi.lib().gb_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """plain `pytest` on a machine without CUDA: the `gpu` tests are skipped, not failed (the product path itself has
    no CPU fallback and raises)"""
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200): run with -m gpu on the GPU box")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
