import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One libsqk context on cuda:0 for the whole GPU session (fails loudly without a GPU)."""
    import squigglekit_b200 as sqk
    c = sqk.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
