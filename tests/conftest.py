import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")

# D2H copies into pageable memory are staged through pinned buffers (hb_api.cu, d2h_copy_2d); with a small staging size
# the tests walk every piece shape -- several rows per piece, one row per piece, parts of one row
os.environ.setdefault("HB_D2H_STAGE_KB", "24")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built():
    """Build (incrementally) the native library and the oracle; tests never run on stale binaries."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
