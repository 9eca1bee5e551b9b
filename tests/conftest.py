import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Build the CUDA C-ABI libraries and the oracle (the checker) once per session."""
    from miluphcuda_b200 import build
    build.build_all()
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import build_oracle
    build_oracle.build()
    yield
