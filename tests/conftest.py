import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rs-vio_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _n_gpus():
    try:
        import rsvio_b200
        return rsvio_b200.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    if _n_gpus() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (GPU parity tests run under gpurun)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def rsvio():
    import rsvio_b200
    rsvio_b200.load_library()
    return rsvio_b200
