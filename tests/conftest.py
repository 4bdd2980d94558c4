import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def O():
    from tests.oracle_lib import oracle

    return oracle()


@pytest.fixture(scope="session")
def ctx():
    """A device context; GPU tests fail loudly (no skip, no fallback) without a device."""
    import ecckd_b200 as E

    c = E.Context(0)
    yield c
    c.close()
