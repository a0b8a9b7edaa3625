import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, 'tests'), os.path.join(REPO, 'tools')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def built():
    """Make sure the CUDA library and the test-only host build exist."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope='session')
def engine(built):
    from nexoclom_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()
