import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fx():
    from is3d_b200 import tables
    return tables.load_fixture()


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checkers (C oracle, and the reference binary where /root/reference exists) and the product library."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    from is3d_b200 import api, build
    if not os.path.exists(api.LIB_PATH):
        build.build()
