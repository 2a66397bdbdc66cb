import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Builds libruvgnn.so and the oracle once per session (no-op when up to date)."""
    import __graft_entry__ as ge

    ge.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from tests.oracle_binding import Oracle

    return Oracle()
