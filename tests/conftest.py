import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """On a box that has a GPU the host-only tests (oracle, packers, hnsw dump reader, store layouts, plan
    model, C / C++ callers) join the `-m gpu` run: the driver's GPU tier then sees every row of the scope
    table, not only the device tests.  Without a GPU nothing changes (`-m "not gpu"` selects them)."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        return
    for item in items:
        if item.get_closest_marker("gpu") is None:
            item.add_marker(pytest.mark.gpu)


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
