import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import __graft_entry__ as ge  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (C-ABI wrapper).  Builds libb200pt.so if it is missing."""
    if not os.path.exists(os.path.join(ge.PKG_DIR, "libb200pt.so")):
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only)."""
    ge.build_oracle()
    import oracle_lib
    return oracle_lib


@pytest.fixture(scope="session")
def gpu(pkg):
    pkg.init(0)
    return pkg
