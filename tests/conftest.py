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


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """The reference-vector comparisons skip while tests/golden/reference_vectors*.txt are absent (they can only be
    generated with cargo, which this image lacks).  Say so loudly at the end of every run instead of hiding it among
    the skips: a green run does NOT mean the CUDA path was compared with the Rust reference itself."""
    golden = os.path.join(ROOT, "tests", "golden")
    missing = [n for n in ("reference_vectors.txt", "reference_vectors_b.txt") if not os.path.exists(os.path.join(golden, n))]
    if missing:
        n = sum(1 for r in terminalreporter.stats.get("skipped", []) if "reference_vectors" in str(getattr(r, "longrepr", "")))
        terminalreporter.write_sep("=", "PARITY AGAINST THE REFERENCE ITSELF: UNPINNED", red=True, bold=True)
        terminalreporter.write_line(
            f"tests/golden/{' and '.join(missing)} absent: {n} reference-vector comparison(s) skipped in this run. The oracle is "
            "pinned by OpenCV / Pillow / integer-tent checks only where an independent implementation exists "
            "(DESIGN.md section 3); generate the vectors with tools/reference_vectors* on a machine with cargo.")
