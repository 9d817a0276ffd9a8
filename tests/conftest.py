import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ge.PKG_DIR, "libdcpgpu.so")):
        ge.build()
    return ge.load_pkg()


@pytest.fixture(scope="session")
def o32():
    import orc
    return orc.Oracle(double=False)


@pytest.fixture(scope="session")
def o64():
    import orc
    return orc.Oracle(double=True)
