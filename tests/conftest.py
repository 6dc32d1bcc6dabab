import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_CAD = "/root/reference/cad_models"     # exists in the build container only


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def have_reference():
    return os.path.isdir(REFERENCE_CAD)


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """Build the oracle and the test-only host simulation (g++); libqspush.so is built by __graft_entry__.build()."""
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "hostsim"), "-s"])
    yield
