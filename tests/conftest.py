import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the C-ABI library and the oracle exist (prebuilt files are reused on the GPU box)."""
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build_lib()
    from oracle import oracle as O
    O.build()
    return True
