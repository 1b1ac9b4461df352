import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Every test needs the oracle library; GPU tests need the product library too. Build once."""
    import __graft_entry__ as g
    g.build(quiet=True)
    # oracle vectors are generated single-threaded: the reference's Green-Gauss boundary loop has a data
    # race on corner cells, and its own regression tests run with OMP_NUM_THREADS=1
    import orc
    orc.set_threads(1)
