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


@pytest.fixture(scope="session", autouse=True)
def _flip_report():
    """BV_WRITE_FLIPS=1: what the parity checks listed as flips goes to gpurun_out/flips_observed.json (see tests/util.py)."""
    yield
    if os.environ.get("BV_WRITE_FLIPS"):
        from tests import util
        util.write_observed_flips(os.path.join(ROOT, "gpurun_out", "flips_observed.json"))
