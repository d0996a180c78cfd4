import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The reference's own C++ compiled into oracle/_ref (test infrastructure)."""
    from oracle import ref as _ref
    if not _ref.available():
        pytest.skip("oracle/_ref/libforte2_ref.so not built")
    return _ref
