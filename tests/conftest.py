import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """The CUDA library, built in-tree (nvcc cross-compiles without a GPU)."""
    from basevar_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import loader
    loader.build_oracle()
    return loader.load_oracle()
